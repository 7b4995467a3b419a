"""One appearance head at the benchmark's size for ncu captures (training-mode forward, stashed backward, inference forward):
    ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 1 -c 1 -o gpurun_out/prof_fwd python scripts/profile_heads.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 892741
torch.manual_seed(0)
head = AppearanceHead(51, (128, 128, 64), 3, 3, False, "gelu", False).cuda()
pos = (torch.rand(n, 51, device="cuda") * 2 - 1)
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
g = torch.randn(n, 3, device="cuda") / n
stash = head.new_stash(n)
flat = torch.zeros(head.num_params(), device="cuda")
dpos = torch.zeros_like(pos)
for _ in range(3):
    out, _ = head.forward_train(pos, dirs, nrm, stash=stash)
    head.backward_into(pos, dirs, nrm, g, flat, dpos, False, None, stash=stash, fwd_out=out)
    head(pos, dirs, nrm)
torch.cuda.synchronize()
