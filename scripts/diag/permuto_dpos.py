"""GPU diagnostic: position gradient of the product vs the reference kernels, level by level (one-hot window)."""
import ctypes, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import permuto as op
import test_gpu_permuto as T

ref = ctypes.CDLL(str(T.REF_PATH))
enc, pos = T._setup(n=30000, seed=3)
n = pos.shape[0]
grad = torch.randn(n, 52, generator=torch.Generator().manual_seed(9)).cuda()
args = (pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(), enc.random_shift_per_level.detach().cpu().numpy())
for lvl in [0, 1, 2, 5, 10, 15, 19, 23]:
    w = torch.zeros(24, device="cuda"); w[lvl] = 1.0
    _, d_pos = enc._launch_backward(enc.lattice_values, pos, w, grad, None, None, want_lattice=False, want_positions=True)
    _, r_pos = T._ref_backward(ref, enc, pos, w, grad)
    _, o_pos = op.backward(*args, w.cpu().numpy(), op.from_rows(grad.cpu().numpy()), fma=True, dtype=np.float64)
    d, r = d_pos.cpu().numpy(), r_pos.cpu().numpy()
    rms = np.sqrt(np.mean(r ** 2))
    bad = np.abs(d - r).max(axis=1) > 1e-5 * rms
    bad_o = np.abs(r - o_pos).max(axis=1) > 1e-5 * rms
    bad_do = np.abs(d - o_pos).max(axis=1) > 1e-5 * rms
    print(f"lvl {lvl}: rms {rms:.3e} ours-vs-ref max {np.abs(d-r).max():.3e} rows differing {bad.sum()}  ref-vs-oracle rows {bad_o.sum()}  ours-vs-oracle rows {bad_do.sum()}")
    if bad.sum():
        i = int(np.argmax(np.abs(d - r).max(axis=1)))
        print("   worst row", i, "ours", d[i], "ref", r[i], "oracle", o_pos[i])
# forward: oracle vs ref per level
theirs = T._ref_forward(ref, enc, pos, enc.anneal_window).cpu().numpy()
want = op.forward(*args, np.ones(24, np.float32), True, 1.0, fma=True)
for lvl in range(0, 26):
    neq = (theirs[lvl] != want[lvl]).any(axis=0)
    print(f"fwd lvl {lvl}: oracle != ref on {neq.sum()} of {n} positions, max abs {np.abs(theirs[lvl]-want[lvl]).max():.3e}")
