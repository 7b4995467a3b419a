"""Times the permutohedral hash encoding kernels alone (CUDA events, L2 flushed between iterations).
    python scripts/bench_permuto.py [n_positions] [reps] [mode: random|rays]
Algorithmic bytes per position: 24 levels x 4 vertices x 8 B gathered (768 B, L2-resident tables) +
12 B position in + 4*out_cols B out; the backward scatters the same 768 B as atomics and reads 4*out_cols B of gradient.
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.encoding import PermutoEncoding  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 892741
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
mode = sys.argv[3] if len(sys.argv) > 3 else "random"
L, cap = 24, 1 << 18
torch.manual_seed(0)
enc = PermutoEncoding(3, cap, L, 2, np.geomspace(1.0, 1e-4, L), True, True, 1.0)
with torch.no_grad():
    enc.lattice_values.copy_(torch.randn(L, cap, 2))
if mode == "rays":   # hit points of neighbouring camera rays on a surface: coherent
    t = torch.arange(n, dtype=torch.float32, device="cuda")
    u, v = (t % 800) / 800.0, torch.div(t, 800, rounding_mode="floor") / (n / 800.0)
    pos = torch.stack([0.2 + 0.6 * u, 0.2 + 0.6 * v, 0.5 + 0.1 * torch.sin(6 * u) * torch.cos(5 * v)], dim=1).contiguous()
else:
    pos = torch.rand(n, 3, device="cuda")
out_cols = 51
grad = torch.randn(n, out_cols, device="cuda")
out = torch.empty(n, out_cols, device="cuda")
d_lat = torch.zeros_like(enc.lattice_values)
d_pos = torch.zeros_like(pos)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tf, tb, tbp = [], [], []
for it in range(reps + 3):
    flush.zero_()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    enc._launch_forward(enc.lattice_values, pos, enc.anneal_window, out_cols, None, None, out=out)
    e[1].record()
    enc._launch_backward(enc.lattice_values, pos, enc.anneal_window, grad, None, None, want_lattice=True, want_positions=False, d_lattice=d_lat)
    e[2].record()
    enc._launch_backward(enc.lattice_values, pos, enc.anneal_window, grad, None, None, want_lattice=True, want_positions=True, d_lattice=d_lat,
                         d_positions=d_pos)
    e[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
        tbp.append(e[2].elapsed_time(e[3]))
med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
gather = n * L * 4 * 8
print(f"permuto n={n} L={L} cap=2^18 {mode}: fwd {med(tf):.4f} ms ({n/med(tf)/1e6:.2f} Gpos/s, gathers {gather/med(tf)/1e6:.0f} GB/s, "
      f"rows out {n*out_cols*4/med(tf)/1e6:.0f} GB/s)  bwd(lattice) {med(tb):.4f} ms ({gather/med(tb)/1e6:.0f} GB/s of atomics)  "
      f"bwd(lattice+pos) {med(tbp):.4f} ms", flush=True)
