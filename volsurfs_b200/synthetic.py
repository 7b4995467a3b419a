"""Synthetic workloads of the BASELINE.json configs.  All tensors are generated on the CPU
in fp32 from a seeded ``torch.Generator`` so that the CPU oracle and the GPU kernels consume identical bits."""
from __future__ import annotations

import math

import numpy as np
import torch

BASE_SEED = 20250328


def gen(seed_offset: int = 0) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(BASE_SEED + seed_offset)
    return g


def dense_layers(n_rays: int, K: int, seed_offset: int = 1, p_hit: float = 0.8, exact_frac: float = 0.01):
    """Config C1-style dense K-layer inputs in MESH order (index 0 = innermost).

    hit ~ Bernoulli(p_hit); alpha ~ U(0,1)*hit with `exact_frac` exact zeros and exact ones among the hits;
    rgb ~ U(0,1); z increasing with the layer index from the camera's point of view means the OUTERMOST layer
    (index K-1) is nearest: z = sorted U(0.5,3.5) descending with layer index.  Upstream grads ~ N(0,1)."""
    g = gen(seed_offset)
    hit = torch.rand(n_rays, K, generator=g) < p_hit
    alpha = torch.rand(n_rays, K, 1, generator=g)
    u = torch.rand(n_rays, K, 1, generator=g)
    alpha = torch.where(u < exact_frac, torch.zeros_like(alpha), alpha)
    alpha = torch.where(u > 1 - exact_frac, torch.ones_like(alpha), alpha)
    alpha = alpha * hit.unsqueeze(-1)
    rgb = torch.rand(n_rays, K, 3, generator=g)
    z = torch.sort(torch.rand(n_rays, K, generator=g) * 3.0 + 0.5, dim=1, descending=True).values.unsqueeze(-1)
    grads = {
        "g_rgb": torch.randn(n_rays, 3, generator=g),
        "g_depth": torch.randn(n_rays, 1, generator=g),
        "g_acc": torch.randn(n_rays, 1, generator=g),
        "g_bgT": torch.randn(n_rays, 1, generator=g),
    }
    return {"hit": hit, "alpha": alpha.contiguous(), "rgb": rgb.contiguous(), "z": z.contiguous(), **grads}


def pack_dense(hit: torch.Tensor, *dense: torch.Tensor):
    """Dense [N,K,d] (mesh order) -> packed [S,d] in outer->inner order + ray_start_end_idx [N,2] int32 with (-1,-1)
    for empty rays.  Host-side data preparation for tests and benchmarks (numpy)."""
    h = hit.numpy()[:, ::-1]
    N, K = h.shape
    cnt = h.sum(axis=1).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
    se = np.stack([start, start + cnt], axis=1).astype(np.int32)
    se[cnt == 0] = -1
    ray, j = np.nonzero(h)
    layer = K - 1 - j
    out = [torch.from_numpy(se)]
    for t in dense:
        a = t.numpy()
        out.append(torch.from_numpy(np.ascontiguousarray(a[ray, layer].reshape(len(ray), -1))))
    return out


def all_hit_packed(n_rays: int, K: int, seed_offset: int = 2):
    """K samples on every ray (the roofline shape of the headline metric: B(s) = 64 + 56*K bytes/ray)."""
    g = gen(seed_offset)
    S = n_rays * K
    start = torch.arange(n_rays, dtype=torch.int32) * K
    se = torch.stack([start, start + K], dim=1).contiguous()
    return {
        "se": se,
        "alpha": torch.rand(S, 1, generator=g),
        "rgb": torch.rand(S, 3, generator=g),
        "z": torch.rand(S, 1, generator=g) * 3 + 0.5,
        "g_rgb": torch.randn(n_rays, 3, generator=g),
        "g_depth": torch.randn(n_rays, 1, generator=g),
        "g_acc": torch.randn(n_rays, 1, generator=g),
        "g_bgT": torch.randn(n_rays, 1, generator=g),
    }


def nerf_packets(n_rays: int, seed_offset: int = 3, max_per_ray: int = 1024, p_empty: float = 0.35, mean: float = 96.0,
                 sigma: float = 0.9):
    """Config C3: per-ray sample counts: p_empty rays with 0 samples, the rest min(max, ceil(LogNormal(ln mean, sigma)));
    sigma_density ~ Exp(20) with 70 % zeros, dt ~ 2/1024; alpha = 1 - exp(-density*dt); x = 1 - alpha + 1e-6."""
    g = gen(seed_offset)
    empty = torch.rand(n_rays, generator=g) < p_empty
    ln = torch.exp(torch.randn(n_rays, generator=g) * sigma + math.log(mean))
    cnt = torch.clamp(torch.ceil(ln), max=max_per_ray).to(torch.int64)
    cnt[empty] = 0
    S = int(cnt.sum())
    start = torch.cumsum(cnt, 0) - cnt
    se = torch.stack([start, start + cnt], dim=1).to(torch.int32)
    se[cnt == 0] = -1
    dens = -torch.log1p(-torch.rand(S, 1, generator=g).clamp(max=1 - 1e-7)) / 20.0 * 1024.0
    dens = dens * (torch.rand(S, 1, generator=g) > 0.7)
    dt = torch.full((S, 1), 2.0 / 1024.0)
    alpha = 1.0 - torch.exp(-dens * dt)
    ray_of = torch.repeat_interleave(torch.arange(n_rays), cnt)
    pos = torch.arange(S) - start[ray_of]
    z = (0.5 + pos.float() * (2.0 / 1024.0)).unsqueeze(1)
    return {
        "se": se.contiguous(), "counts": cnt, "alpha": alpha.contiguous(), "x": (1 - alpha + 1e-6).contiguous(),
        "rgb": torch.rand(S, 3, generator=g), "z": z.contiguous(), "dt": dt,
        "g_rgb": torch.randn(n_rays, 3, generator=g), "g_depth": torch.randn(n_rays, 1, generator=g),
        "g_acc": torch.randn(n_rays, 1, generator=g), "g_bgT": torch.randn(n_rays, 1, generator=g),
    }


def composite_bytes(n_rays: int, n_samples: int) -> int:
    """Algorithmic HBM bytes of fused compositing fwd+bwd: sum_r (64 + 56 s_r)."""
    return 64 * n_rays + 56 * n_samples
