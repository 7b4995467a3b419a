"""Permutohedral encoder kernels on the benchmark's real hit points (5 shells, 800x800 view), in the packed (ray-major) order and in a
layer-major order of the same points: how much of the lattice backward is the order the atomics arrive in.
    python scripts/bench_permuto_hits.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.encoding import PermutoHashEncoder  # noqa: E402
from volsurfs_b200.pipeline import make_synthetic_renderer  # noqa: E402
from volsurfs_b200.raytracer import pack_layer_hits  # noqa: E402
from volsurfs_b200.synthetic import camera_rays  # noqa: E402

dev = torch.device("cuda", 0)
renderer, _ = make_synthetic_renderer(K=5, hidden=(128, 128, 64), pos_dim=51)
o, d = camera_rays(800, 800, azimuth_deg=30.0)
rec = renderer.tracer.trace_layers(o.to(dev), d.to(dev))
rsp = pack_layer_hits(rec["rays_o"], rec["rays_d"], rec["depth"], rec["tri"], rec["u"], rec["v"], t_far=renderer.tracer.t_far, exact_size=True)
n = int(rsp.total_dev.item())
pos = rsp.samples_3d[:n].contiguous()
layer = rsp.samples_layer[:n].long()
perm = torch.sort(layer, stable=True).indices
orders = {"ray-major (packed)": pos, "layer-major": pos[perm].contiguous(), "shuffled": pos[torch.randperm(n, device=dev)].contiguous()}
torch.manual_seed(0)
enc = PermutoHashEncoder(bb_sides=2.0, device=dev)
e = enc.encoder
with torch.no_grad():
    e.lattice_values.normal_(0.0, 0.1)
grad = torch.randn(n, 51, device=dev)
out = torch.empty(n, 51, device=dev)
d_lat = torch.zeros_like(e.lattice_values)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, p in orders.items():
    tf, tb = [], []
    for it in range(13):
        flush.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        e._launch_forward(e.lattice_values, p, enc.window(None), 51, enc.bb_sides, None, out=out)
        ev[1].record()
        e._launch_backward(e.lattice_values, p, enc.window(None), grad, enc.bb_sides, None, want_lattice=True, d_lattice=d_lat)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 3:
            tf.append(ev[0].elapsed_time(ev[1]))
            tb.append(ev[1].elapsed_time(ev[2]))
    med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
    print(f"{name:22s} n={n}: fwd {med(tf):.4f} ms  bwd(lattice) {med(tb):.4f} ms", flush=True)
key = rsp.samples_layer[:n].contiguous()
tb = []
for it in range(13):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    e._launch_backward(e.lattice_values, pos, enc.window(None), grad, enc.bb_sides, None, want_lattice=True, d_lattice=d_lat, order_key=key)
    ev[1].record()
    torch.cuda.synchronize()
    if it >= 3:
        tb.append(ev[0].elapsed_time(ev[1]))
print(f"{'ray-major, key = layer':22s} n={n}: bwd(lattice) {med(tb):.4f} ms", flush=True)
