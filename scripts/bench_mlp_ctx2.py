"""Stashed backward timed on the real step's tensors, then with single inputs swapped for synthetic ones."""
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
head = renderer.rgb_head
stash = renderer._stash_rgb
rgb, d_rgb = out["samples_rgb"], out["d_rgb"]
flat = torch.zeros(head.num_params(), device="cuda")
dpos = torch.zeros_like(feats)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = int(rsp.total_dev.item())
print("hits", n, "d_rgb absmax", float(d_rgb[:n].abs().max()), "zeros frac", float((d_rgb[:n] == 0).float().mean()))


def t(tag, dirs, nrm, gg, fo, nv, st=stash):
    tb = []
    for it in range(10):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        e[0].record()
        head.backward_into(feats, dirs, nrm, gg, flat, dpos, False, nv, stash=st, fwd_out=fo)
        e[1].record()
        torch.cuda.synchronize()
        tb.append(e[0].elapsed_time(e[1]))
    print(f"{tag}: {sorted(tb)[5]:.4f} ms", flush=True)


t("real", rsp.samples_dirs, rsp.samples_normals, d_rgb, rgb, rsp.total_dev)
t("g -> randn*1e-6", rsp.samples_dirs, rsp.samples_normals, torch.randn_like(d_rgb) * 1e-6, rgb, rsp.total_dev)
t("fwd_out -> rand", rsp.samples_dirs, rsp.samples_normals, d_rgb, torch.rand_like(rgb), rsp.total_dev)
st2 = head.new_stash(N * 5)
o2, _ = head.forward_train(feats, rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev, stash=st2)
t("fresh stash same inputs", rsp.samples_dirs, rsp.samples_normals, d_rgb, o2, rsp.total_dev, st2)
rd = torch.nn.functional.normalize(torch.randn_like(rsp.samples_dirs), dim=1)
o3, _ = head.forward_train(feats, rd, rd, n_valid_dev=rsp.total_dev, stash=st2)
t("stash from random dirs", rd, rd, d_rgb, o3, rsp.total_dev, st2)
t("stash from random dirs + g randn", rd, rd, torch.randn_like(d_rgb) * 1e-6, o3, rsp.total_dev, st2)
