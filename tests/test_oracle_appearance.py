"""CPU: the appearance oracle reproduces vectors produced by the reference's own MLP / SHEncoder classes and alpha-decay lines."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import appearance as oa


@pytest.mark.parametrize("name", ["appearance_rgb_128", "appearance_alpha_64"])
def test_head_matches_reference_classes(name):
    g = np.load(GOLDEN / f"{name}.npz")
    n_layers = len(g["hidden"]) + 1
    Ws = [torch.from_numpy(g[f"W{i}"]).requires_grad_(True) for i in range(n_layers)]
    bs = [torch.from_numpy(g[f"b{i}"]).requires_grad_(True) for i in range(n_layers)]
    pos = torch.from_numpy(g["pos"]).requires_grad_(True)
    dirs, normals = torch.from_numpy(g["dirs"]), torch.from_numpy(g["normals"])
    assert np.array_equal(oa.sh_encode(dirs, 3).numpy(), g["sh"])
    out = oa.head_forward(pos, dirs, normals, Ws, bs, normal_dep=bool(g["normal_dep"]))
    assert np.allclose(out.detach().numpy(), g["out"], rtol=0, atol=1e-7)
    assert np.allclose(oa.alpha_decay(out[:, :1], dirs, normals).detach().numpy(), g["alpha_decayed"], rtol=0, atol=1e-7)
    (out * torch.from_numpy(g["g_out"])).sum().backward()
    assert np.allclose(pos.grad.numpy(), g["d_pos"], rtol=1e-5, atol=1e-8)
    for i in range(n_layers):
        assert np.allclose(Ws[i].grad.numpy(), g[f"dW{i}"], rtol=1e-4, atol=1e-7)
        assert np.allclose(bs[i].grad.numpy(), g[f"db{i}"], rtol=1e-4, atol=1e-7)
