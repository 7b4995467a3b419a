import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from volsurfs_b200.synthetic import nerf_packets
from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = nerf_packets(640000, seed_offset=3)
rsp = RaySamplesPacked(0, 0, 0, 1)
rsp.ray_start_end_idx = d["se"].cuda()
a, c, z = d["alpha"].cuda(), d["rgb"].cuda(), d["z"].cuda()
g = [d[k].cuda() for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")]
for _ in range(3):
    VR.composite(rsp, a, c, z, mode=mode)
    VR.composite_backward(rsp, a, c, z, *g, mode=mode)
torch.cuda.synchronize()
