"""time the K-layer intersector on the C2 geometry (800x800 rays, 5 shells x ~100k triangles):
    VS_TRACE_VARIANT=0 python scripts/bench_trace.py"""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.raytracer import ShellTracer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays, shell_meshes  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5
tracer = ShellTracer(shell_meshes(K=K))
for shuffle in (None, 3):
    o, d = camera_rays(800, 800, shuffle_seed=shuffle)
    o, d = o.cuda(), d.cuda()
    for _ in range(3):
        rec = tracer.trace_layers(o, d)
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rec = tracer.trace_layers(o, d)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    hits = int((rec["depth"] <= 100).sum())
    print(json.dumps({"variant": os.environ.get("VS_TRACE_VARIANT", "0"), "K": K, "shuffled": shuffle is not None, "ms_median": round(ts[10], 4),
                      "ms_min": round(ts[0], 4), "mray_layer_pairs_s": round(o.shape[0] * K / ts[10] / 1e3, 1), "hits": hits,
                      "overflow": tracer.overflowed()}), flush=True)
