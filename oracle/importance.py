"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement (numpy fp32 scalars, sequential per ray) of the importance-sampling chain of the reference:

* pcg32 .......................... kernels/volsurfs/pcg32.h:32-34 (defaults), :60-70 (next_uint), :84-95 (next_float), :158-180 (advance)
* map_range_val .................. kernels/volsurfs/VolumeRenderingGPU.cuh:15-21
* binary_search .................. kernels/volsurfs/VolumeRenderingGPU.cuh:481-505
* importance_sample_gpu .......... kernels/volsurfs/VolumeRenderingGPU.cuh:507-678 (host: src/VolumeRendering.cu:466-548)
* combine_ray_samples_packets_gpu  kernels/volsurfs/VolumeRenderingGPU.cuh:680-894 (host: src/VolumeRendering.cu:550-669)

Both functions return the UNCOMPACTED buffers exactly as the reference kernels leave them (constructor fill -1 where nothing is
written); compaction is oracle/packing.py.  Pin: tests/test_gpu_reference_kernels.py runs the reference's own kernels
(oracle/_ref/libvolsurfs_ref.so, compiled from the reference sources) on the same inputs on the GPU box and compares both this
restatement and the product against them; the reference ships no vectors for these functions.  Arithmetic is fp32 without FMA
contraction — nvcc contracts a*b+c in the reference build, so values may differ from the reference kernels in the last ulp.

Deviations (documented in the tests): a 1-sample segment makes the reference's binary search spin forever — here it returns the
only index; a packet with no samples for a ray makes the reference's merge read row ``start-1`` — here it counts as exhausted.
"""
from __future__ import annotations

import numpy as np

F = np.float32
M64 = (1 << 64) - 1
PCG_MULT = 0x5851F42D4C957F2D
PCG_DEFAULT_STATE = 0x853C49E6748FEA9B
PCG_DEFAULT_INC = 0xDA3E39CB94B95BDB


class Pcg32:
    def __init__(self, state=PCG_DEFAULT_STATE, inc=PCG_DEFAULT_INC):
        self.state, self.inc = state, inc

    def next_uint(self) -> int:
        old = self.state
        self.state = (old * PCG_MULT + self.inc) & M64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def next_float(self) -> np.float32:
        u = (self.next_uint() >> 9) | 0x3F800000
        return F(np.array([u], dtype=np.uint32).view(np.float32)[0] - F(1.0))

    def advance(self, delta: int = 1 << 32) -> None:
        cur_mult, cur_plus, acc_mult, acc_plus = PCG_MULT, self.inc, 1, 0
        delta &= M64
        while delta > 0:
            if delta & 1:
                acc_mult = (acc_mult * cur_mult) & M64
                acc_plus = (acc_plus * cur_mult + cur_plus) & M64
            cur_plus = ((cur_mult + 1) * cur_plus) & M64
            cur_mult = (cur_mult * cur_mult) & M64
            delta >>= 1
        self.state = (acc_mult * self.state + acc_plus) & M64

    def copy(self) -> "Pcg32":
        return Pcg32(self.state, self.inc)


def map_range_val(v, in_start, in_end, out_start, out_end):
    v, in_start, in_end, out_start, out_end = F(v), F(in_start), F(in_end), F(out_start), F(out_end)
    clamped = max(in_start, min(in_end, v))
    if in_start >= in_end:
        return out_end
    return F(out_start + F(F(F(out_end - out_start) / F(in_end - in_start)) * F(clamped - in_start)))


def binary_search(cdf, val, imin, imax):
    if imax <= imin:
        return imax
    while imax >= imin:
        imid = imin + (imax - imin) // 2
        if cdf[imid] > val:
            imax = imid
        else:
            imin = imid
        if imax - imin == 1:
            return imax
    return imax


def importance_sample(rays_o, rays_d, se, z, cdf, n_imp, jitter=False, rng: Pcg32 | None = None):
    """-> dict(samples_3d [N*n_imp,3], samples_dirs, samples_z [N*n_imp,1], ray_start_end_idx [N,2]) with -1 where untouched"""
    n_rays = se.shape[0]
    z = np.asarray(z, F).reshape(-1)
    cdf = np.asarray(cdf, F).reshape(-1)
    out_3d = np.full((n_rays * n_imp, 3), -1, F)
    out_dirs = np.full((n_rays * n_imp, 3), -1, F)
    out_z = np.full((n_rays * n_imp, 1), -1, F)
    out_se = np.full((n_rays, 2), -1, np.int32)
    rng = rng or Pcg32()
    dist = F(1.0 / (n_imp + 1))
    mov = F(float(dist) / 2.0)
    lo, hi = F(0.0 + 1e-6), F(1.0 - 1e-6)
    for r in range(n_rays):
        start, end = int(se[r, 0]), int(se[r, 1])
        if end - start == 0:
            continue
        g = rng.copy()  # passed by value to the kernel: every thread starts from the same state
        o, d = rays_o[r].astype(F), rays_d[r].astype(F)
        for i in range(n_imp):
            u = F(dist + F(F(i) * dist))
            if jitter:
                g.advance(r)
                u = F(u + map_range_val(g.next_float(), 0.0, 1.0, -mov, mov))
            u = max(lo, min(hi, u))
            imax = binary_search(cdf, u, start, end - 1)
            imin = max(imax - 1, 0)
            z_imp = map_range_val(u, cdf[imin], cdf[imax], z[imin], z[imax])
            row = r * n_imp + i
            out_3d[row] = o + z_imp * d
            out_dirs[row] = d
            out_z[row, 0] = z_imp
        out_se[r] = (r * n_imp, r * n_imp + n_imp)
    return {"samples_3d": out_3d, "samples_dirs": out_dirs, "samples_z": out_z, "ray_start_end_idx": out_se}


def combine_ray_samples_packets(se1, idx1, p1, d1, z1, v1, se2, idx2, p2, d2, z2, v2, min_dist):
    """-> dict of the uncompacted combined buffers (size n1+n2, -1 / arange fill like the RaySamplesPacked constructor)"""
    n_rays = se1.shape[0]
    n1, n2 = z1.shape[0], z2.shape[0]
    vd = v1.shape[1]
    z1f, z2f = np.asarray(z1, F).reshape(-1), np.asarray(z2, F).reshape(-1)
    c1 = (se1[:, 1] - se1[:, 0]).astype(np.int64)
    c2 = (se2[:, 1] - se2[:, 0]).astype(np.int64)
    out_start = np.concatenate([[0], np.cumsum(c1 + c2)[:-1]]).astype(np.int32)  # VolumeRendering.cu:595-603
    n = n1 + n2
    c_idx = np.arange(n, dtype=np.int32).reshape(-1, 1)
    c_3d = np.full((n, 3), -1, F)
    c_dirs = np.full((n, 3), -1, F)
    c_z = np.full((n, 1), -1, F)
    c_val = np.full((n, vd), -1, F)
    c_se = np.full((n_rays, 2), -1, np.int32)
    md = F(min_dist)
    for r in range(n_rays):
        s1, s2 = int(se1[r, 0]), int(se2[r, 0])
        m1, m2 = int(c1[r]), int(c2[r])
        if m1 == 0 and m2 == 0:
            continue
        a = b = written = 0
        done1, done2 = m1 == 0, m2 == 0
        prec = F(0.0)
        base = int(out_start[r])
        for _ in range(m1 + m2):
            if done1 and done2:
                break
            za = F(1e10) if done1 else z1f[s1 + a]
            zb = F(1e10) if done2 else z2f[s2 + b]
            take1 = za < zb
            zz = za if take1 else zb
            if not (F(zz - prec) < md):
                dst = base + written
                if take1:
                    src = s1 + a
                    c_idx[dst], c_3d[dst], c_dirs[dst], c_val[dst] = idx1[src], p1[src], d1[src], v1[src]
                else:
                    src = s2 + b
                    c_idx[dst], c_3d[dst], c_dirs[dst], c_val[dst] = idx2[src], p2[src], d2[src], v2[src]
                c_z[dst, 0] = zz
                prec = zz
                written += 1
            if take1:
                if a + 1 >= m1:
                    done1 = True
                else:
                    a += 1
            else:
                if b + 1 >= m2:
                    done2 = True
                else:
                    b += 1
        c_se[r] = (base, base + written)
    return {"samples_idx": c_idx, "samples_3d": c_3d, "samples_dirs": c_dirs, "samples_z": c_z, "samples_values": c_val,
            "ray_start_end_idx": c_se, "out_start": out_start}
