"""TEST INFRASTRUCTURE: the whole legacy-head hot path restated on the CPU from the oracle pieces — reference-faithful BVH trace
(oracle/raytrace_oracle.c, the reference kernel's arithmetic) -> packing (oracle/packing.py) -> torch heads (oracle/appearance.py; the
reference's MLP / SHEncoder semantics) -> dense K-layer compositing of volsurfs.py:601-640,708 in packed form -> L1 loss
(utils/losses.py:14-19) -> torch autograd.  Used by tests/test_gpu_pipeline.py and by __graft_entry__.smoke()."""
from __future__ import annotations

import numpy as np
import torch

from oracle import appearance as oa
from oracle.packing import pack_layer_hits
from oracle.raytrace import OracleRayTracer


def head_params(head, dtype=torch.float32):
    return ([l.weight.detach().cpu().to(dtype).requires_grad_(True) for l in head.layers],
            [l.bias.detach().cpu().to(dtype).requires_grad_(True) for l in head.layers])


def packed_composite_torch(se, alpha, rgb, bg=1.0):
    """rgb_fg + bgT * bg for packed samples, per-ray exclusive cumprod, differentiable (torch); se [N,2] numpy"""
    N = se.shape[0]
    cnt = np.maximum(se[:, 1] - se[:, 0], 0)
    kmax = int(cnt.max()) if N else 0
    # scatter to dense [N, kmax] (missing samples: alpha 0, as the dense reference path does for misses, volsurfs.py:456)
    ray = np.repeat(np.arange(N), cnt)
    pos = np.arange(int(cnt.sum())) - np.repeat(np.where(cnt > 0, se[:, 0], 0), cnt)
    a = torch.zeros(N, kmax, 1, dtype=alpha.dtype)
    c = torch.zeros(N, kmax, 3, dtype=alpha.dtype)
    idx = (torch.from_numpy(ray), torch.from_numpy(pos))
    a = a.index_put(idx, alpha)
    c = c.index_put(idx, rgb)
    Tc = torch.cumprod(1 - a, dim=1)
    T = torch.cat([torch.ones_like(Tc[:, :1]), Tc[:, :-1]], dim=1)
    w = T * a
    fg = (c * w).sum(dim=1)
    bgT = Tc[:, -1] if kmax else torch.ones(N, 1, dtype=alpha.dtype)
    return fg + bgT * bg, w


def oracle_step(meshes, rays_o, rays_d, feats, rgb_head, alpha_head, gt, dtype=torch.float32):
    """rays / feats / gt: CPU tensors; heads: AppearanceHead modules (weights read, not modified).  Returns a dict with the image, the
    loss, the packed reference packet and autograd gradients of both heads' parameters (flat, AppearanceHead.split_flat layout) and of
    their positional features."""
    o, d = rays_o.numpy(), rays_d.numpy()
    lay = OracleRayTracer(meshes, contract="device").trace_layers(o, d, mode="bvh")
    unc, layer_of_slot = pack_layer_hits(o, d, lay["is_hit"].T, lay["depth"].T)
    want = unc.compact_to_valid_samples()
    S = want.get_total_nr_samples()
    K = len(meshes)
    slot = want.samples_idx[:, 0]
    ray, layer = slot // K, layer_of_slot[slot]
    normals = torch.from_numpy(np.stack([r["normals"] for r in lay["per_mesh"]])[layer, ray]).to(dtype)
    dirs = torch.from_numpy(want.samples_dirs).to(dtype)
    f_rgb = feats[:S].to(dtype).clone().requires_grad_(True)
    f_alpha = feats[:S].to(dtype).clone().requires_grad_(True)
    Wr, br = head_params(rgb_head, dtype)
    Wa, ba = head_params(alpha_head, dtype)
    rgb = oa.head_forward(f_rgb, dirs, normals, Wr, br)
    alpha = oa.alpha_decay(oa.head_forward(f_alpha, dirs, normals, Wa, ba), dirs, normals)
    pred, _ = packed_composite_torch(want.ray_start_end_idx, alpha, rgb)
    loss = (pred - gt.to(dtype)).abs().mean()
    loss.backward()

    def flat(Ws, bs):
        return torch.cat([p.grad.reshape(-1) for W, b in zip(Ws, bs) for p in (W, b)])

    return {"rgb": pred.detach(), "loss": loss.detach(), "packet": want, "n_samples": S, "samples_rgb": rgb.detach(), "samples_alpha": alpha.detach(),
            "grad_rgb": flat(Wr, br), "grad_alpha": flat(Wa, ba), "d_features_rgb": f_rgb.grad, "d_features_alpha": f_alpha.grad}
