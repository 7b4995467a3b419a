"""One forward + one backward compositing launch per workload, for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:composite -o gpurun_out/prof python scripts/profile_composite.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.synthetic import all_hit_packed, nerf_packets  # noqa: E402
from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR  # noqa: E402

modes = [int(m) for m in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1]
which = sys.argv[2] if len(sys.argv) > 2 else "shells"
# shells: 2^22 rays x 5; shells24: the roofline size of SURVEY 8d (2^24 rays x 5, 5.8 GB of traffic per fwd+bwd pair); nerf: 200 k rays of
# config[2]-shaped packets; nerf640k: BASELINE config[2] itself (640 k rays, ~59 M samples)
d = {"shells": lambda: all_hit_packed(1 << 22, 5), "shells24": lambda: all_hit_packed(1 << 24, 5),
     "nerf": lambda: nerf_packets(200000, seed_offset=3), "nerf640k": lambda: nerf_packets(640000, seed_offset=3)}[which]()
rsp = RaySamplesPacked(0, 0, 0, 1)
rsp.ray_start_end_idx = d["se"].cuda()
a, c, z = d["alpha"].cuda(), d["rgb"].cuda(), d["z"].cuda()
g = [d[k].cuda() for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")]
for mode in modes:
    for _ in range(2):
        VR.composite(rsp, a, c, z, mode=mode)
        VR.composite_backward(rsp, a, c, z, *g, mode=mode)
torch.cuda.synchronize()
