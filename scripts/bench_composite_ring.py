"""A/B of the long-ray compositing kernels on BASELINE config C3 (640k rays, <= 1024 samples per ray) and a shorter-ray packet:
scan family (mode 2), coarsened scan (mode 3) and the cp.async ring family (mode 8).  python scripts/bench_composite_ring.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from bench_composite import run  # noqa: E402
from volsurfs_b200.synthetic import all_hit_packed, nerf_packets  # noqa: E402

torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run("c3_nerf_packets", nerf_packets(640000, seed_offset=3), [2, 3, 8], flush)
run("nerf_mean24", nerf_packets(640000, seed_offset=5, max_per_ray=128, mean=24.0), [2, 3, 8], flush)
run("shells_allhit_K9", all_hit_packed(1 << 21, 9), [1, 2, 8], None)
