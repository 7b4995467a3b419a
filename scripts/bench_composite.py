"""A/B micro-benchmark of the compositing kernel families (run on the GPU box):
    python scripts/bench_composite.py [--rays 16777216] [--K 5]
Prints one JSON line per (workload, mode): Mrays/s, GB/s of algorithmic traffic, fraction of the measured HBM peak."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.synthetic import all_hit_packed, composite_bytes, dense_layers, nerf_packets, pack_dense  # noqa: E402
from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR  # noqa: E402


def peak_gbs():
    p = Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_kernel(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def run(name, d, modes, flush):
    n_rays, S = d["se"].shape[0], d["alpha"].shape[0]
    rsp = RaySamplesPacked(0, 0, 0, 1)
    rsp.ray_start_end_idx = d["se"].cuda()
    a, c, z = d["alpha"].cuda(), d["rgb"].cuda(), d["z"].cuda()
    g = [d[k].cuda() for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")]
    peak, how = peak_gbs()
    nbytes = composite_bytes(n_rays, S)
    for mode in modes:
        try:
            tf, tf_min = time_kernel(lambda: VR.composite(rsp, a, c, z, mode=mode), flush=flush)
            tb, tb_min = time_kernel(lambda: VR.composite_backward(rsp, a, c, z, *g, mode=mode), flush=flush)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"workload": name, "mode": mode, "error": str(exc)}))
            continue
        t = tf + tb
        print(json.dumps({
            "workload": name, "mode": mode, "n_rays": n_rays, "n_samples": S, "fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
            "mrays_s": round(n_rays / t / 1e3, 1), "gbs": round(nbytes / t / 1e6, 1), "frac_of_hbm_peak": round(nbytes / t / 1e6 / peak, 4),
            "fwd_gbs": round((8 * n_rays + 20 * S + 24 * n_rays) / tf / 1e6, 1), "bwd_gbs": round((32 * n_rays + 36 * S) / tb / 1e6, 1),
            "peak": how,
        }), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1 << 24)
    ap.add_argument("--K", type=int, default=5)
    ap.add_argument("--nerf-rays", type=int, default=640000)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    run(f"shells_allhit_K{args.K}", all_hit_packed(args.rays, args.K), [1, 4, 3], None)  # 5.8 GB of traffic >> L2
    run("shells_allhit_K9", all_hit_packed(args.rays // 2, 9), [1, 2, 3], None)
    d = dense_layers(1 << 22, 5, seed_offset=1)
    se, a, c, z = pack_dense(d["hit"], d["alpha"], d["rgb"], d["z"])
    d.update(se=se, alpha=a, rgb=c, z=z)
    run("shells_bernoulli0.8_K5", d, [1, 4, 3], flush)
    run("c2_800x800_K5", all_hit_packed(640000, 5), [1, 4], flush)
    run("c3_nerf_packets", nerf_packets(args.nerf_rays, seed_offset=3), [2, 5, 6, 7, 3], flush)
    run("nerf_mean24", nerf_packets(args.nerf_rays, seed_offset=5, max_per_ray=128, mean=24.0), [2, 5, 6, 3], flush)
