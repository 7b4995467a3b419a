"""Why does the stashed backward take longer inside the step than alone?  Times it under controlled variations."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402

n = 892741
torch.manual_seed(0)
head = AppearanceHead(51, (128, 128, 64), 3, 3, False, "gelu", False).cuda()


def run(tag, cap, use_nvalid, gscale, sparse, reps=12):
    pos = (torch.rand(cap, 51, device="cuda") * 2 - 1)
    dirs = torch.nn.functional.normalize(torch.randn(cap, 3, device="cuda"), dim=1)
    nrm = torch.nn.functional.normalize(torch.randn(cap, 3, device="cuda"), dim=1)
    g = torch.randn(cap, 3, device="cuda") * gscale
    if sparse:
        g = g * (torch.rand(cap, 1, device="cuda") < 0.3)
    nv = torch.tensor([n], dtype=torch.int64, device="cuda") if use_nvalid else None
    stash = head.new_stash(cap)
    flat = torch.zeros(head.num_params(), device="cuda")
    dpos = torch.zeros_like(pos)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tb = []
    for it in range(reps + 3):
        out, _ = head.forward_train(pos, dirs, nrm, n_valid_dev=nv, stash=stash)
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        head.backward_into(pos, dirs, nrm, g, flat, dpos, False, nv, stash=stash, fwd_out=out)
        e[1].record()
        torch.cuda.synchronize()
        if it >= 3:
            tb.append(e[0].elapsed_time(e[1]))
    print(f"{tag}: bwd {sorted(tb)[len(tb)//2]:.4f} ms", flush=True)


run("cap=n, no n_valid, g~1e-6", n, False, 1e-6, False)
run("cap=n, n_valid, g~1e-6", n, True, 1e-6, False)
run("cap=3.2M, n_valid, g~1e-6", 3200000, True, 1e-6, False)
run("cap=3.2M, n_valid, g~1e-6 sparse", 3200000, True, 1e-6, True)
run("cap=3.2M, n_valid, g~1 ", 3200000, True, 1.0, False)
