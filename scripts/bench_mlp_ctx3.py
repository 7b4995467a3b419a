"""Stashed backward inside a continuously running loop vs. after a host sync."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.pipeline import make_synthetic_renderer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays  # noqa: E402

renderer, _ = make_synthetic_renderer(K=5)
o, d = camera_rays(800, 800)
o, d = o.cuda(), d.cuda()
N = o.shape[0]
g = torch.Generator().manual_seed(100)
feats = (torch.rand(N * 5, 51, generator=g) * 2 - 1).cuda()
gt = torch.rand(N, 3, generator=g).cuda()
out = renderer.render_fwd_bwd(o, d, feats, gt)
rsp = out["ray_samples_packed"]
rgb, alpha, d_rgb, d_alpha = out["samples_rgb"], out["samples_alpha"], out["d_rgb"], out["d_alpha"]
hr, ha = renderer.rgb_head, renderer.alpha_head
sr, sa = renderer._stash_rgb, renderer._stash_alpha
fr = torch.zeros(hr.num_params(), device="cuda")
fa = torch.zeros(ha.num_params(), device="cuda")
dpr, dpa = torch.zeros_like(feats), torch.zeros_like(feats)
dirs, nrm, nv = rsp.samples_dirs, rsp.samples_normals, rsp.total_dev


def loop(tag, pre, sync, reps=20):
    tb, ta, tf = [], [], []
    for it in range(reps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        if pre == "fwd":
            hr.forward_train(feats, dirs, nrm, n_valid_dev=nv, stash=sr)
            ha.forward_train(feats, dirs, nrm, n_valid_dev=nv, stash=sa)
        elif pre == "trace":
            renderer.tracer.trace_layers(o, d)
        e[1].record()
        if sync:
            torch.cuda.synchronize()
        e[2].record()
        hr.backward_into(feats, dirs, nrm, d_rgb, fr, dpr, False, nv, stash=sr, fwd_out=rgb)
        e[3].record()
        ha.backward_into(feats, dirs, nrm, d_alpha, fa, dpa, False, nv, stash=sa, fwd_out=alpha)
        e[4].record()
        if sync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    # only the last iteration's events are kept on purpose (steady state)
    print(f"{tag}: pre {e[0].elapsed_time(e[1]):.4f}  bwd_rgb {e[2].elapsed_time(e[3]):.4f}  bwd_alpha {e[3].elapsed_time(e[4]):.4f} ms", flush=True)


loop("pre=none sync", None, True)
loop("pre=none nosync", None, False)
loop("pre=fwd sync", "fwd", True)
loop("pre=fwd nosync", "fwd", False)
loop("pre=trace nosync", "trace", False)
loop("pre=none nosync (again)", None, False)
