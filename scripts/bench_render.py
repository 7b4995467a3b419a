"""Inference render of the K-layer path (BASELINE config[4]: DTU-scale 1600x1200 image, 9 nested layers; config[1]: 800x800, 5 layers):
trace + pack + normals + two legacy heads (forward only) + compositing, rows of the image sharded over the ranks when launched with
torchrun.  CUDA events between barriers, max over ranks; rank 0 prints one JSON line.
    python scripts/bench_render.py [--height 1200] [--width 1600] [--layers 9] [--steps 10]"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200 import _lib  # noqa: E402
from volsurfs_b200.dist import shard_rays  # noqa: E402
from volsurfs_b200.pipeline import make_synthetic_renderer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--layers", type=int, default=9)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    offset = 0.006 if args.layers >= 9 else 0.01   # r_k = r0 + k * 0.006 for the 9-layer stress config
    renderer, _ = make_synthetic_renderer(K=args.layers, offset=offset)
    o, d = camera_rays(args.height, args.width)
    o, d = shard_rays(o, d, rank, world)           # contiguous row blocks
    o, d = o.to(dev), d.to(dev)
    N = o.shape[0]
    feats = torch.rand(N * args.layers, 51, device=dev) * 2 - 1   # the positional encoder's output (synthetic, as in bench.py)
    with torch.no_grad():
        for _ in range(3):
            out = renderer.render(o, d, feats)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        before = _lib.lib().vs_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = renderer.render(o, d, feats)
        e1.record()
        torch.cuda.synchronize()
    launches = _lib.lib().vs_launch_count() - before
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    hits = torch.tensor([int(out["ray_samples_packed"].total_dev.item())], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(hits)
    if rank == 0:
        total = args.height * args.width
        print(json.dumps({"workload": f"{args.layers}-layer render, {args.width}x{args.height} image, legacy heads [128,128,64]", "n_gpus": world,
                          "rays": total, "hits": int(hits[0]), "ms_per_frame": round(float(ms[0]), 3),
                          "mrays_s": round(total / float(ms[0]) / 1e3, 1), "kernels_per_frame": launches // args.steps}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
