"""CUDA-event timing of one appearance head at the benchmark's size (training forward, stashed backward):
    python scripts/time_heads.py [n_samples] [iters]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 892741
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
head = AppearanceHead(51, (128, 128, 64), 3, 3, False, "gelu", False).cuda()
pos = torch.rand(n, 51, device="cuda") * 2 - 1
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
g = torch.randn(n, 3, device="cuda") / n
stash = head.new_stash(n)
flat = torch.zeros(head.num_params(), device="cuda")
dpos = torch.zeros_like(pos)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = None


def fwd():
    global out
    out, _ = head.forward_train(pos, dirs, nrm, stash=stash)


def bwd():
    head.backward_into(pos, dirs, nrm, g, flat, dpos, False, None, stash=stash, fwd_out=out)


for _ in range(3):
    fwd()
    bwd()
print(f"n={n} fwd_train {timed(fwd):.4f} ms  bwd_stashed {timed(bwd):.4f} ms")
