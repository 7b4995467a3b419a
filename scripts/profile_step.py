"""Three passes of the full hot path (bench.py's step) for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:mlp_fwd -s 2 -c 1 -o gpurun_out/prof_mlp python scripts/profile_step.py
    ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.pipeline import make_synthetic_renderer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
renderer, _ = make_synthetic_renderer(K=5)
o, d = camera_rays(800, 800)
o, d = o.cuda(), d.cuda()
N = o.shape[0]
g = torch.Generator().manual_seed(100)
feats = (torch.rand(N * 5, 51, generator=g) * 2 - 1).cuda()
gt = torch.rand(N, 3, generator=g).cuda()
for _ in range(n_steps):
    renderer.render_fwd_bwd(o, d, feats, gt)
torch.cuda.synchronize()
