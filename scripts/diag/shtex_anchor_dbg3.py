import sys, ctypes
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_shtex import _case, _module_from
d, nets, C, deg = _case("shtex_rgb_anchor")
m = _module_from(d, nets, C, deg)
net = m.neural_textures[2].model
uv = torch.from_numpy(d["uv"]).cuda()
feat = net.encode(uv, 0, True, [16, 16])
r = 117
print("uv", uv[r].tolist(), "feat row", feat[r].tolist())
W0 = net.weights[0].detach().half().float()
x0 = feat[r] @ W0.t()
x0d = feat[r].double() @ W0.double().t()
idx = x0.abs().argsort()[:6]
print("smallest |pre-act| layer0:", [(int(i), float(x0[i]), float(x0d[i])) for i in idx])
h0 = torch.relu(x0).half().float()
W1 = net.weights[1].detach().half().float()
x1 = h0 @ W1.t(); x1d = h0.double() @ W1.double().t()
idx = x1.abs().argsort()[:6]
print("smallest |pre-act| layer1:", [(int(i), float(x1[i]), float(x1d[i])) for i in idx])
print("any duplicate rows of feat 117:", [int(i) for i in torch.nonzero((feat == feat[r]).all(1)).flatten()])
