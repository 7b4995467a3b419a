"""One training step of the full path with the reference's DEFAULT appearance (per-layer SH neural textures for colour and transparency,
config/volsurfs/base_5.cfg) at BASELINE config[1] size: trace + pack + uv + 2 x K SHNeuralTextures + compositing, forward and backward
(torch autograd drives the backward kernels).  CUDA events, median.
    python scripts/bench_textured_step.py [H] [W] [K] [reps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200 import _lib  # noqa: E402
from volsurfs_b200.pipeline import make_synthetic_textured_renderer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 800
W = int(sys.argv[2]) if len(sys.argv) > 2 else 800
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
renderer, _ = make_synthetic_textured_renderer(K=K, table_init=0.5)
o, d = camera_rays(H, W)
o, d = o.cuda(), d.cuda()
gt = torch.rand(H * W, 3, device="cuda")
ts, tf = [], []
for it in range(reps + 2):
    for p in renderer.parameters():
        p.grad = None
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    before = _lib.lib().vs_launch_count()
    e0.record()
    out = renderer.render(o, d)
    loss = (out["rgb"] - gt).abs().mean()
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    launches = _lib.lib().vs_launch_count() - before
    if it >= 2:
        tf.append(e0.elapsed_time(e1))
        ts.append(e0.elapsed_time(e2))
med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
S = out["ray_samples_packed"].get_total_nr_samples()
print(f"textured step {H}x{W} rays x {K} shells, {S} hits, {2 * K} SHNeuralTextures ({8 * K} texture networks): forward {med(tf):.2f} ms, "
      f"forward+backward {med(ts):.2f} ms = {H * W / med(ts) / 1e3:.2f} Mrays/s; {launches} kernels of this library per step", flush=True)
