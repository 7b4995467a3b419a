import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from bench_composite import run
from volsurfs_b200.synthetic import all_hit_packed, nerf_packets
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run("c3_nerf_packets", nerf_packets(640000, seed_offset=3), [8], flush)
run("nerf_mean24", nerf_packets(640000, seed_offset=5, max_per_ray=128, mean=24.0), [1, 8], flush)
run("nerf_mean400", nerf_packets(200000, seed_offset=6, max_per_ray=1024, mean=400.0, sigma=0.5, p_empty=0.1), [2, 8], flush)
