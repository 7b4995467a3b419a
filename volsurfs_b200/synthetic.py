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
                 sigma: float = 0.9, min_per_ray: int = 1):
    """Config C3: per-ray sample counts: p_empty rays with 0 samples, the rest min(max, ceil(LogNormal(ln mean, sigma)));
    sigma_density ~ Exp(20) with 70 % zeros, dt ~ 2/1024; alpha = 1 - exp(-density*dt); x = 1 - alpha + 1e-6."""
    g = gen(seed_offset)
    empty = torch.rand(n_rays, generator=g) < p_empty
    ln = torch.exp(torch.randn(n_rays, generator=g) * sigma + math.log(mean))
    cnt = torch.clamp(torch.ceil(ln), min=min_per_ray, max=max_per_ray).to(torch.int64)
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


# ---------------------------------------------------------------------------------------------------------------------
# geometry: nested "kitten-like" shells and pinhole camera rays (config C2 / C5)
# ---------------------------------------------------------------------------------------------------------------------
def shell_meshes(K: int = 5, n_lat: int = 224, n_lon: int = 224, r_base: float = 0.30, offset: float = 0.01, seed_offset: int = 40):
    """K nested star-shaped shells: UV-sphere topology (2*n_lon*(n_lat-1) triangles, ~100k at 224x224), radius field
    r0(theta,phi) = r_base*(1 + 0.25*sum_m a_m f_m(theta,phi)) with fixed-seed low-frequency lobes, shell k at
    r0 + k*offset (k = 0 innermost ... K-1 outermost), as produced by offset-SDF level sets
    (reference utils/mesh_extraction.py:375-405).  Returns a list of (vertices [V,3] f32, faces [F,3] i32)."""
    rng = np.random.default_rng(BASE_SEED + seed_offset)
    amps = rng.uniform(-1, 1, 4)
    phases = rng.uniform(0, 2 * np.pi, 4)
    theta = np.linspace(0.0, np.pi, n_lat + 1)[1:-1]                 # interior latitudes
    phi = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    T, P = np.meshgrid(theta, phi, indexing="ij")

    def lobes(t, p):
        return (amps[0] * np.sin(t) * np.cos(p + phases[0]) + amps[1] * np.sin(2 * t) * np.sin(2 * p + phases[1])
                + amps[2] * np.sin(t) ** 2 * np.cos(3 * p + phases[2]) + amps[3] * np.cos(2 * t + phases[3]))

    r_grid = r_base * (1 + 0.25 * 0.5 * lobes(T, P))
    r_north = r_base * (1 + 0.25 * 0.5 * lobes(0.0, 0.0))
    r_south = r_base * (1 + 0.25 * 0.5 * lobes(np.pi, 0.0))
    dirs = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], -1).reshape(-1, 3)
    rows, cols = n_lat - 1, n_lon
    idx = np.arange(rows * cols).reshape(rows, cols)
    nxt = np.roll(idx, -1, axis=1)
    quads_a = np.stack([idx[:-1], idx[1:], nxt[1:]], -1).reshape(-1, 3)
    quads_b = np.stack([idx[:-1], nxt[1:], nxt[:-1]], -1).reshape(-1, 3)
    north, south = rows * cols, rows * cols + 1
    cap_n = np.stack([np.full(cols, north), idx[0], nxt[0]], -1)
    cap_s = np.stack([np.full(cols, south), nxt[-1], idx[-1]], -1)
    faces = np.concatenate([quads_a, quads_b, cap_n, cap_s]).astype(np.int32)
    meshes = []
    for k in range(K):
        rk = (r_grid + k * offset).reshape(-1, 1)
        v = np.concatenate([dirs * rk, [[0, 0, r_north + k * offset]], [[0, 0, -(r_south + k * offset)]]]).astype(np.float32)
        meshes.append((np.ascontiguousarray(v), faces.copy()))
    return meshes


def shell_face_uvs(n_lat: int = 224, n_lon: int = 224) -> np.ndarray:
    """Per-face-vertex texture coordinates [F,3,2] for the faces of ``shell_meshes`` (what ``TensorMesh.get_faces_uvs()`` returns in the
    reference, volsurfs.py:511): the lat/lon chart u = phi / 2pi, v = theta / pi, with the seam column unwrapped (u = 1 instead of 0) and
    the poles pinned to the u of the cap triangle's first ring vertex."""
    theta = np.linspace(0.0, np.pi, n_lat + 1)[1:-1]
    rows, cols = n_lat - 1, n_lon
    v_ring = (theta / np.pi).astype(np.float64)
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")

    def uv(rr, cc):  # cc may be == cols (the unwrapped seam)
        return np.stack([cc / cols, v_ring[rr]], -1)

    qa = np.stack([uv(r[:-1], c[:-1]), uv(r[1:], c[1:]), uv(r[1:], c[1:] + 1)], -2).reshape(-1, 3, 2)
    qb = np.stack([uv(r[:-1], c[:-1]), uv(r[1:], c[1:] + 1), uv(r[:-1], c[:-1] + 1)], -2).reshape(-1, 3, 2)
    cc = np.arange(cols)
    cap_n = np.stack([np.stack([cc / cols, np.zeros(cols)], -1), uv(np.zeros(cols, int), cc), uv(np.zeros(cols, int), cc + 1)], -2)
    cap_s = np.stack([np.stack([cc / cols, np.ones(cols)], -1), uv(np.full(cols, rows - 1), cc + 1), uv(np.full(cols, rows - 1), cc)], -2)
    return np.ascontiguousarray(np.concatenate([qa, qb, cap_n, cap_s]).astype(np.float32))


def camera_rays(height: int = 800, width: int = 800, fov_deg: float = 40.0, radius: float = 1.5, azimuth_deg: float = 30.0,
                elevation_deg: float = 20.0, shuffle_seed=None):
    """Pinhole camera on an orbit looking at the origin; rays in scanline order, directions normalised
    (mvdatasets/utils/raycasting.py:167-247 conventions).  Returns (rays_o [H*W,3], rays_d [H*W,3]) float32 tensors."""
    az, el = np.deg2rad(azimuth_deg), np.deg2rad(elevation_deg)
    eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, [0, 0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    f = 0.5 * width / np.tan(0.5 * np.deg2rad(fov_deg))
    j, i = np.meshgrid(np.arange(height) + 0.5, np.arange(width) + 0.5, indexing="ij")
    d = ((i - 0.5 * width)[..., None] * right + (0.5 * height - j)[..., None] * up + f * fwd).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(eye, d.shape)
    if shuffle_seed is not None:
        perm = np.random.default_rng(shuffle_seed).permutation(d.shape[0])
        d = d[perm]
    return torch.from_numpy(np.ascontiguousarray(o, dtype=np.float32)), torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
