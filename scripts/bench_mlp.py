"""Times the appearance-head kernels alone (CUDA events, L2 flushed by the working set): training-mode forward + stashed backward.
    python scripts/bench_mlp.py [n_samples] [hidden e.g. 128,128,64] [reps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 892741
hidden = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (128, 128, 64)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
torch.manual_seed(0)
for out_dim, decay in ((3, False), (1, True)):
    head = AppearanceHead(51, hidden, out_dim, 3, False, "gelu", decay).cuda()
    pos = (torch.rand(n, 51, device="cuda") * 2 - 1)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
    nrm = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
    g = torch.randn(n, out_dim, device="cuda") / n
    stash = head.new_stash(n)
    flat = torch.zeros(head.num_params(), device="cuda")
    dpos = torch.zeros_like(pos)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tf, tb, ti = [], [], []
    for it in range(reps + 3):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        out, _ = head.forward_train(pos, dirs, nrm, stash=stash)
        e[1].record()
        head.backward_into(pos, dirs, nrm, g, flat, dpos, False, None, stash=stash, fwd_out=out)
        e[2].record()
        head(pos, dirs, nrm)
        e[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(e[0].elapsed_time(e[1]))
            tb.append(e[1].elapsed_time(e[2]))
            ti.append(e[2].elapsed_time(e[3]))
    med = lambda v: sorted(v)[len(v) // 2]
    flops = 2 * sum(a * b for a, b in zip(head.dims[:-1], head.dims[1:])) * n
    print(f"head out={out_dim} hidden={hidden} n={n}: fwd(train) {med(tf):.4f} ms ({flops/med(tf)/1e9:.1f} TF/s)  bwd(stashed) {med(tb):.4f} ms "
          f"({2*flops/med(tb)/1e9:.1f} TF/s)  fwd(inference) {med(ti):.4f} ms  stash {stash.numel()/1e6:.0f} MB", flush=True)
