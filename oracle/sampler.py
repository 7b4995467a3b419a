"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement (numpy fp32 scalars, one Python loop per ray: small cases only) of the reference's ray samplers and the
occupancy-grid queries under them:

* pos_to_lin_idx / morton3D / distance_to_next_voxel ... kernels/volsurfs/occ_grid_helpers.h:13-33,55-79,126-190
* compute_samples_fg .................................. kernels/volsurfs/RaySamplerGPU.cuh:141-271 (+ compaction, src/RaySampler.cu:236)
* compute_samples_fg_in_grid_occupied_regions .......... RaySamplerGPU.cuh:273-488
* compute_samples_bg .................................. RaySamplerGPU.cuh:39-139
* contract_samples / uncontract_samples ............... RaySamplerGPU.cuh:528-658 (before update_dt, src/RaySampler.cu:378,425)
* get_rays_t_near_t_far / check_occupancy ............. kernels/volsurfs/OccupancyGridGPU.cuh:318-441
* pcg32 ............................................... oracle/importance.py:Pcg32 (kernels/volsurfs/pcg32.h)

PIN STATUS.  The reference has no tests or vectors for these functions, but its kernels compile here from the sources where they lie
(oracle/ref_sampler_harness.cu -> oracle/_ref/libsampler_ref.so; Eigen::Vector3f replaced by a 3-float stand-in), so this restatement
AND the product kernels are pinned bit-exact against the reference kernels themselves on the GPU (tests/test_gpu_sampler.py).
``fma=True`` (default) reproduces the contractions nvcc applies to the reference source: ``ray_o + t * ray_d``, ``t + c * rnd`` and
helper_math's ``lerp`` become fused multiply-adds (the product of two fp32 values is exact in fp64, so fp32(fp64(a)*fp64(b)+fp64(c))
is an FMA up to a double rounding that needs a 29-bit tie; none occurs in the fixtures).
"""
from __future__ import annotations

import numpy as np

from .importance import Pcg32

F = np.float32
EPS = F(1e-6)


def _mad(a, b, c, fma=True):
    if fma:
        return F(np.float64(a) * np.float64(b) + np.float64(c))
    return F(F(a * b) + c)


def _to_u32(x) -> int:
    """float -> uint32_t as cvt.rzi.u32.f32 does it: truncate, negatives / NaN -> 0, saturate"""
    if not np.isfinite(x):
        return 0 if (np.isnan(x) or x < 0) else 0xFFFFFFFF
    if x <= 0:
        return 0
    return min(int(x), 0xFFFFFFFF)


def spread_bits(v: int) -> int:
    """occ_grid_helpers.h:13-22, truncated to 32 bits like ``uint32_t xx = expand_bits(x)`` (:28-30)"""
    w = v & 0x00000000001FFFFF
    w = (w | w << 32) & 0x001F00000000FFFF
    w = (w | w << 16) & 0x001F0000FF0000FF
    w = (w | w << 8) & 0x010F00F00F00F00F
    w = (w | w << 4) & 0x10C30C30C30C30C3
    w = (w | w << 2) & 0x1249249249249249
    return w & 0xFFFFFFFF


def morton3d(x: int, y: int, z: int) -> int:
    return (spread_bits(x) | (spread_bits(y) << 1) | (spread_bits(z) << 2)) & 0xFFFFFFFF


def pos_to_lin_idx(pos, n: int, extent) -> int:
    """occ_grid_helpers.h:55-79 -> signed int like ``int morton_idx``"""
    c = [_to_u32(F(F(F(pos[a] / F(extent[a])) + F(0.5)) * F(n))) for a in range(3)]
    m = morton3d(*c)
    return m - (1 << 32) if m >= (1 << 31) else m


def distance_to_next_voxel(pos, d, n: int, extent):
    """occ_grid_helpers.h:126-190: distance ALONG THE AXIS to the next voxel boundary (not divided by the direction), + 1e-6"""
    if abs(d[0]) < EPS and abs(d[1]) < EPS and abs(d[2]) < EPS:
        return F(1e10)
    t3 = []
    for a in range(3):
        t = F(1e10)
        if abs(d[a]) > EPS:
            q = F(F(pos[a] / F(extent[a])) * F(n))
            sgn = F(1.0) if d[a] > 0 else (F(-1.0) if d[a] < 0 else F(0.0))
            prime = F(np.floor(F(q + sgn)))
            t = F(F(abs(F(prime - q)) / F(n)) * F(extent[a]))
        t3.append(t)
    return F(min(min(t3[0], t3[1]), t3[2]) + EPS)


class Grid:
    def __init__(self, n, extent, occupancy, roi, values=None):
        self.n, self.extent = int(n), [F(e) for e in extent]
        self.occ, self.roi = np.asarray(occupancy, bool), np.asarray(roi, bool)
        self.values = None if values is None else np.asarray(values, np.float32)

    def inside(self, v):
        return 0 <= v < self.n ** 3

    def occupied(self, v):
        return bool(self.roi[v] and self.occ[v])


def _pos(o, d, t, fma):
    return [_mad(t, d[a], o[a], fma) for a in range(3)]


def _clamp(v, a, b):
    return F(max(a, min(b, v)))


def samples_fg(rays_o, rays_d, t_entry, t_exit, min_dist, min_nr, max_nr, jitter=False, rng: Pcg32 | None = None, grid: Grid | None = None,
               fma=True):
    """UNCOMPACTED result of compute_samples_fg (grid=None) / compute_samples_fg_in_grid_occupied_regions: dict of numpy arrays laid out
    like the reference's packet (ray r owns slots [r*max_nr, (r+1)*max_nr)); constructor fills as in src/RaySamplesPacked.cu:13-48."""
    rays_o, rays_d = np.asarray(rays_o, np.float32), np.asarray(rays_d, np.float32)
    t_entry, t_exit = np.asarray(t_entry, np.float32).reshape(-1), np.asarray(t_exit, np.float32).reshape(-1)
    N, cap = rays_o.shape[0], rays_o.shape[0] * max_nr
    out = dict(samples_idx=np.arange(cap, dtype=np.int32).reshape(-1, 1), samples_3d=np.full((cap, 3), -1, np.float32),
               samples_dirs=np.full((cap, 3), -1, np.float32), samples_z=np.full((cap, 1), -1, np.float32),
               samples_dt=np.full((cap, 1), -1, np.float32), ray_start_end_idx=np.full((N, 2), -1, np.int32),
               ray_max_dt=np.full((N, 1), -1, np.float32))
    min_dist = F(min_dist)
    rng = rng if rng is not None else Pcg32()
    for r in range(N):
        o, d, t_start, t_end = rays_o[r], rays_d[r], t_entry[r], t_exit[r]
        if grid is None:
            dist = F(t_end - t_start)
        else:
            dist, step, t = F(0), F(0), t_start
            while t < t_end:
                pos = _pos(o, d, t, fma)
                v = pos_to_lin_idx(pos, grid.n, grid.extent)
                if not grid.inside(v):
                    break
                if grid.occupied(v):
                    dist = F(dist + step)
                step = distance_to_next_voxel(pos, d, grid.n, grid.extent)
                t = F(t + step)
            dist = _clamp(dist, F(0), F(t_end - t_start))
        to_create, spacing = 0, F(0)
        if dist > 0:
            if dist > min_dist:
                to_create = int(F(dist / min_dist))
                to_create = max(0, min(to_create, max_nr))
                spacing = F(dist / F(to_create))
            else:
                to_create, spacing = 1, dist
        created, slot0 = 0, r * max_nr
        if to_create > 0 and to_create >= min_nr:
            t, to_next = t_start, F(0)
            if jitter:
                g = rng.copy()
                g.advance(r)
                rnd = g.next_float()
                if grid is None:
                    t = _mad(spacing, rnd, t, fma)
                else:
                    to_next = F(spacing * rnd)
            while t < t_end:
                t = _clamp(t, t_start, t_end)
                pos = _pos(o, d, t, fma)
                if created >= to_create:
                    break
                emit, occ = True, True
                if grid is not None:
                    v = pos_to_lin_idx(pos, grid.n, grid.extent)
                    if not grid.inside(v):
                        break
                    occ = grid.occupied(v)
                    emit = occ and to_next == 0
                if emit:
                    s = slot0 + created
                    out["samples_3d"][s] = pos
                    out["samples_dirs"][s] = d
                    out["samples_z"][s, 0] = t
                    created += 1
                    if grid is not None:
                        to_next = spacing
                if grid is None:
                    t = F(t + spacing)
                else:
                    to_voxel = distance_to_next_voxel(pos, d, grid.n, grid.extent)
                    if occ:
                        step = F(min(to_voxel, to_next))
                        to_next = F(to_next - step)
                        if to_next <= EPS:
                            to_next = F(0)
                    else:
                        step = to_voxel
                    t = F(t + step)
        if created < min_nr:
            created = 0
        else:
            out["ray_max_dt"][r, 0] = spacing
            out["ray_start_end_idx"][r] = (slot0, slot0 + created)
        out["samples_idx"][slot0 + created:slot0 + max_nr, 0] = -1
    return out


def compact(unc):
    """compact_to_valid_samples (src/RaySamplesPacked.cu:188-273) on the dict of ``samples_fg``"""
    se = unc["ray_start_end_idx"]
    cnt = (se[:, 1] - se[:, 0]).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]]) if len(cnt) else np.zeros(0, np.int64)
    total = int(cnt.sum())
    out = dict(samples_idx=np.zeros((total, 1), np.int32), samples_3d=np.zeros((total, 3), np.float32), samples_dirs=np.zeros((total, 3), np.float32),
               samples_z=np.zeros((total, 1), np.float32), samples_dt=np.zeros((total, 1), np.float32),
               ray_start_end_idx=np.full(se.shape, -1, np.int32), ray_max_dt=unc["ray_max_dt"].copy())
    for r in np.nonzero(cnt > 0)[0]:
        si, so, c = int(se[r, 0]), int(start[r]), int(cnt[r])
        for k in ("samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt"):
            out[k][so:so + c] = unc[k][si:si + c]
        out["ray_start_end_idx"][r] = (so, so + c)
    return out


def samples_bg(rays_o, rays_d, t_start, t_far, nr, jitter=False, rng: Pcg32 | None = None, fma=True):
    """compute_samples_bg (RaySamplerGPU.cuh:39-139): nr samples per ray, uniform in inverse depth"""
    rays_o, rays_d = np.asarray(rays_o, np.float32), np.asarray(rays_d, np.float32)
    t_start = np.asarray(t_start, np.float32).reshape(-1)
    N = rays_o.shape[0]
    out = dict(samples_3d=np.zeros((N * nr, 3), np.float32), samples_dirs=np.zeros((N * nr, 3), np.float32), samples_z=np.zeros((N * nr, 1), np.float32),
               ray_start_end_idx=np.zeros((N, 2), np.int32), ray_max_dt=np.zeros((N, 1), np.float32))
    rng = rng if rng is not None else Pcg32()
    t_far = F(t_far)
    delta_s = F(1.0 / float(nr - 1))
    for r in range(N):
        o, d, ts = rays_o[r], rays_d[r], t_start[r]
        g = rng.copy()
        max_dt, s, t_prec = F(0), F(1), ts
        for i in range(nr):
            t = F(1.0 / float(F(s + EPS)) - 1.0)  # double literals in the reference: evaluated in double, rounded once
            t = F(t + ts)
            t = _clamp(t, ts, t_far)
            if jitter and i != 0 and i != nr - 1:
                g.advance(r)
                interp = g.next_float()
                t = _mad(interp, F(t - t_prec), t_prec, fma)
            k = r * nr + i
            out["samples_z"][k, 0] = t
            out["samples_3d"][k] = _pos(o, d, t, fma)
            out["samples_dirs"][k] = d
            s = F(s - delta_s)
            max_dt = F(max(max_dt, F(t - t_prec)))
            t_prec = t
        out["ray_max_dt"][r, 0] = max_dt
        out["ray_start_end_idx"][r] = (r * nr, r * nr + nr)
    return out


def rays_t_near_t_far(rays_o, rays_d, t_entry, t_exit, grid: Grid, fma=True):
    """OccupancyGridGPU.cuh:318-395"""
    rays_o, rays_d = np.asarray(rays_o, np.float32), np.asarray(rays_d, np.float32)
    t_entry, t_exit = np.asarray(t_entry, np.float32).reshape(-1), np.asarray(t_exit, np.float32).reshape(-1)
    N = rays_o.shape[0]
    near, far = np.zeros((N, 1), np.float32), np.zeros((N, 1), np.float32)
    for r in range(N):
        o, d, ts, te = rays_o[r], rays_d[r], t_entry[r], t_exit[r]
        nr_, fr_, t, first = ts, ts, ts, True
        while t < te:
            pos = _pos(o, d, t, fma)
            v = pos_to_lin_idx(pos, grid.n, grid.extent)
            if not grid.inside(v):
                break
            occ = grid.occupied(v)
            if occ and first:
                nr_, first = t, False
            t = F(t + distance_to_next_voxel(pos, d, grid.n, grid.extent))
            if occ:
                fr_ = _clamp(t, ts, te)
        near[r, 0], far[r, 0] = nr_, fr_
    return near, far


def check_occupancy(points, grid: Grid):
    """OccupancyGridGPU.cuh:397-441"""
    points = np.asarray(points, np.float32)
    occ, val = np.zeros((points.shape[0], 1), bool), np.zeros((points.shape[0], 1), np.float32)
    for i, p in enumerate(points):
        v = pos_to_lin_idx(p, grid.n, grid.extent)
        if grid.inside(v):
            occ[i, 0] = grid.occupied(v)
            val[i, 0] = grid.values[v]
    return occ, val


def _length3(x, y, z):
    """helper_math length() = sqrtf(dot) with the dot product x*x + y*y + z*z as nvcc contracts it in the reference build (read off the SASS of
    oracle/_ref/libsampler_ref.so: FMUL y*y, FFMA x, FFMA z): fma(z,z, fma(x,x, y*y)).  fp64 product + one rounding per fma: products of
    fp32 values are exact in fp64"""
    x, y, z = (np.asarray(v, np.float32).astype(np.float64) for v in (x, y, z))
    t = (y * y).astype(np.float32).astype(np.float64)
    t = (x * x + t).astype(np.float32).astype(np.float64)
    t = (z * z + t).astype(np.float32)
    return np.sqrt(t, dtype=np.float32)


def contract_samples(ray_o, ray_start_end_idx, samples_3d, samples_z, uncontract=False):
    """contract_samples_gpu / uncontract_samples_gpu (RaySamplerGPU.cuh:528-592, 594-658): points with |2x| > 1 are mapped to
    (2 - 1/|2x|) x / |2x| (inverse: 1 / (2 - |2x|)) and their depth re-measured from the ray origin; the rest is copied.
    Returns (samples_3d, samples_z) before the closing update_dt (src/RaySampler.cu:378,425)."""
    ray_o = np.asarray(ray_o, np.float32)
    se = np.asarray(ray_start_end_idx, np.int32)
    p = np.array(samples_3d, np.float32, copy=True)
    z = np.array(samples_z, np.float32, copy=True).reshape(-1, 1)
    if p.shape[0] == 0:
        return p, z
    ray_of = np.repeat(np.arange(se.shape[0]), np.maximum(se[:, 1] - se[:, 0], 0))
    idx = np.concatenate([np.arange(a, b) for a, b in se if b > a]) if ray_of.size else np.zeros(0, np.int64)
    q = p[idx]
    cam = ray_o[ray_of]
    two = F(2.0)
    norm = _length3(q[:, 0] * two, q[:, 1] * two, q[:, 2] * two)
    m = norm > F(1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        factor = (F(1.0) / (two - norm)) if uncontract else (two - F(1.0) / norm)
        qn = ((factor[:, None] * q) / norm[:, None]).astype(np.float32)
    zn = _length3(qn[:, 0] - cam[:, 0], qn[:, 1] - cam[:, 1], qn[:, 2] - cam[:, 2])
    p[idx[m]] = qn[m]
    z[idx[m], 0] = zn[m]
    return p, z


# ---- occupancy-grid maintenance (OccupancyGridGPU.cuh:31-218, src/OccupancyGrid.cu:206-347,446-503) ------------------------------------
def compact_bits(x):
    """morton3D_invert (occ_grid_helpers.h:45-53), vectorised"""
    x = np.asarray(x, np.uint32) & np.uint32(0x49249249)
    x = (x | (x >> np.uint32(2))) & np.uint32(0xC30C30C3)
    x = (x | (x >> np.uint32(4))) & np.uint32(0x0F00F00F)
    x = (x | (x >> np.uint32(8))) & np.uint32(0xFF0000FF)
    x = (x | (x >> np.uint32(16))) & np.uint32(0x0000FFFF)
    return x


def _voxel_axis(c, n, extent, centre_grid, centre_of_voxel):
    """one axis of lin_idx_to_3D (occ_grid_helpers.h:74-113)"""
    x = (c.astype(np.float32) / F(n)).astype(np.float32)
    if centre_grid:
        x = (x - F(0.5)).astype(np.float32)
    if centre_of_voxel:
        x = (x + F(F(1.0 / float(n)) / F(2))).astype(np.float32)
    return (x * F(extent)).astype(np.float32)


def grid_points(point_indices, n, extent, centre=True, jitter=False, rng: Pcg32 | None = None):
    """get_grid_lower_left_voxels_vertices_gpu (centre=False) / get_grid_samples_gpu (centre=True): voxel indices (Morton) -> [P,3].
    Jitter: thread i advances a copy of the generator by 3 i and draws x, y, z; mov = fma(voxel_size, rand, -half_voxel_size)."""
    v = np.asarray(point_indices, np.int32).astype(np.uint32)
    out = np.stack([_voxel_axis(compact_bits(v >> np.uint32(a)), n, extent[a], True, centre) for a in range(3)], axis=1)
    if centre and jitter:
        rng = rng if rng is not None else Pcg32()
        size = [F(F(extent[a]) / F(n)) for a in range(3)]
        for i in range(v.shape[0]):
            g = rng.copy()
            g.advance(3 * i)
            for a in range(3):
                out[i, a] = F(out[i, a] + _mad(size[a], g.next_float(), F(-(size[a] / F(2)))))
    return out


def update_grid_values(point_indices, values, decay, grid_values):
    """update_grid_values_gpu (OccupancyGridGPU.cuh:122-147) for UNIQUE indices (duplicates race in the reference); returns the new grid"""
    g = np.array(grid_values, np.float32, copy=True)
    idx = np.asarray(point_indices, np.int64)
    g[idx] = np.maximum(np.asarray(values, np.float32).reshape(-1), (g[idx] * F(decay)).astype(np.float32))
    return g


def update_grid_occupancy_density(point_indices, n, extent, thresh, check_neighbours, grid_values, occupancy):
    """update_grid_occupancy_with_density_values_gpu (OccupancyGridGPU.cuh:149-218); the neighbourhood is addressed through
    lin_idx_to_3D(.., centre_grid=false, ..) * n, i.e. integer voxel coordinates times the extent, as the reference does"""
    occ = np.array(occupancy, bool, copy=True)
    g = np.asarray(grid_values, np.float32)
    thresh = F(thresh)
    for v in np.asarray(point_indices, np.int64):
        if not check_neighbours:
            occ[v] = not (g[v] <= thresh)
            continue
        p = [F(_voxel_axis(compact_bits(np.uint32(v) >> np.uint32(a)), n, extent[a], False, False) * F(n)) for a in range(3)]
        empty = True
        for i in (-1, 0, 1):
            qx = F(p[0] + F(i))
            if qx < 0 or qx > n - 1:
                continue
            for j in (-1, 0, 1):
                qy = F(p[1] + F(j))
                if qy < 0 or qy > n - 1:
                    continue
                for k in (-1, 0, 1):
                    qz = F(p[2] + F(k))
                    if qz < 0 or qz > n - 1:
                        continue
                    empty = empty and bool(g[morton3d(int(qx), int(qy), int(qz))] <= thresh)
        occ[v] = not empty
    return occ


def update_grid_occupancy_sdf(point_indices, n, extent, logistic_beta, thresh, grid_values, occupancy, return_weight=False):
    """update_grid_occupancy_with_sdf_values_gpu (OccupancyGridGPU.cuh:220-316): occupied when the logistic density
    beta e / (1 + e)^2, e = exp(-beta d), at d = max(|sdf| - half the voxel diagonal, 0) exceeds the threshold.  numpy's float32 exp may
    differ from CUDA's expf in the last bit, so the decision can differ only for weights within an ulp of the threshold (the GPU test
    excludes those; the product itself is compared bit for bit with the reference kernel)."""
    occ = np.array(occupancy, bool, copy=True)
    idx = np.asarray(point_indices, np.int64)
    g = np.asarray(grid_values, np.float32)
    beta = np.asarray(logistic_beta, np.float32).reshape(-1)
    s = [F(F(extent[a]) / F(n)) for a in range(3)]
    diagonal = np.sqrt(F(F(F(s[0] * s[0]) + F(s[1] * s[1])) + F(s[2] * s[2])), dtype=np.float32)
    d = np.clip((np.abs(g[idx]) - F(diagonal / F(2))).astype(np.float32), F(0), F(1e10))
    with np.errstate(over="ignore"):
        e = np.clip(np.exp((-beta * d).astype(np.float32), dtype=np.float32), F(-1e6), F(1e6))
        one_e = (F(1) + e).astype(np.float32)
        w = ((beta * e).astype(np.float32) / (one_e * one_e).astype(np.float32)).astype(np.float32)
    occ[idx] = w > F(thresh)
    return (occ, w) if return_weight else occ
