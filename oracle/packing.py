"""ORACLE (test infrastructure only — never imported by the product).

numpy restatement of the reference's ``RaySamplesPacked`` container and of the
packing step.

* container ctor ............ src/RaySamplesPacked.cu:13-48
* compact_to_valid_samples .. src/RaySamplesPacked.cu:188-273 +
                              kernels/volsurfs/RaySamplesPackedGPU.cuh:172-257
* update_dt ................. kernels/volsurfs/RaySamplesPackedGPU.cuh:14-88
* uncompacted slot layout ... kernels/volsurfs/RaySamplerGPU.cuh:206-271
                              (ray r owns slots [r*M, r*M+cnt_r))
* K-layer hit bookkeeping ... volsurfs_py/methods/volsurfs.py:476-518,601-603

All index arithmetic is exact integer work: the CUDA product must match
bit-for-bit.
"""
from __future__ import annotations

import numpy as np


class RaySamplesPackedNP:
    """Field-for-field numpy twin of RaySamplesPacked (RaySamplesPacked.cuh:7-81)."""

    def __init__(self, nr_rays, max_nr_samples, first_sample_idx=0, values_dim=1):
        f = np.float32
        self.samples_idx = np.arange(first_sample_idx, first_sample_idx + max_nr_samples, dtype=np.int32).reshape(-1, 1)
        self.samples_3d = np.full((max_nr_samples, 3), -1, f)
        self.samples_dirs = np.full((max_nr_samples, 3), -1, f)
        self.samples_z = np.full((max_nr_samples, 1), -1, f)
        self.samples_dt = np.full((max_nr_samples, 1), -1, f)
        self.samples_values = np.full((max_nr_samples, values_dim), -1, f)
        self.ray_start_end_idx = np.full((nr_rays, 2), -1, np.int32)
        self.ray_o = np.full((nr_rays, 3), -1, f)
        self.ray_d = np.full((nr_rays, 3), -1, f)
        self.ray_enter = np.full((nr_rays, 1), -1, f)
        self.ray_exit = np.full((nr_rays, 1), -1, f)
        self.ray_max_dt = np.full((nr_rays, 1), -1, f)
        self.has_samples_values = False
        self.has_dt = False
        self.is_compacted = True

    def get_nr_rays(self):
        return self.ray_start_end_idx.shape[0]

    def get_max_nr_samples(self):
        return self.samples_idx.shape[0]

    def get_nr_samples_per_ray(self):
        return self.ray_start_end_idx[:, 1] - self.ray_start_end_idx[:, 0]

    def get_total_nr_samples(self):
        return int(self.get_nr_samples_per_ray().sum())

    def is_empty(self):
        return self.get_nr_rays() == 0 or self.get_total_nr_samples() == 0

    def compact_to_valid_samples(self):
        """RaySamplesPacked.cu:188-273: exclusive cumsum of the per-ray counts
        (int32), each ray's segment gathered to its offset; empty rays keep
        (-1,-1); samples_idx carries the SOURCE slot index."""
        n_rays = self.get_nr_rays()
        total = self.get_total_nr_samples()
        out = RaySamplesPackedNP(n_rays, total, 0, self.samples_values.shape[1])
        out.ray_o = self.ray_o.copy()
        out.ray_d = self.ray_d.copy()
        out.ray_enter = self.ray_enter.copy()
        out.ray_exit = self.ray_exit.copy()
        out.ray_max_dt = self.ray_max_dt.copy()
        out.has_samples_values = self.has_samples_values
        out.has_dt = self.has_dt
        out.is_compacted = True
        if self.is_empty():
            return out
        cnt = self.get_nr_samples_per_ray().astype(np.int64)
        start_in = self.ray_start_end_idx[:, 0].astype(np.int64)
        start_out = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        for r in np.nonzero(cnt > 0)[0]:
            si, so, c = start_in[r], start_out[r], cnt[r]
            for name in ("samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt", "samples_values"):
                getattr(out, name)[so:so + c] = getattr(self, name)[si:si + c]
            out.ray_start_end_idx[r] = (so, so + c)
        return out

    def update_dt(self, is_background):
        """RaySamplesPackedGPU.cuh:14-88."""
        se = self.ray_start_end_idx
        for r in range(self.get_nr_rays()):
            s, e = int(se[r, 0]), int(se[r, 1])
            n = e - s
            if n == 0:
                continue
            max_dt = self.ray_max_dt[r, 0]
            z = self.samples_z[s:e, 0]
            if n >= 2:
                self.samples_dt[s:e - 1, 0] = np.maximum(np.float32(0), np.minimum(z[1:] - z[:-1], max_dt))  # clamp = fmaxf(a, fminf(f, b))
            if is_background:
                self.samples_dt[e - 1, 0] = np.float32(1e10)
            else:
                self.samples_dt[e - 1, 0] = np.maximum(np.float32(0), np.minimum(self.ray_exit[r, 0] - z[-1], max_dt))
        self.has_dt = True


def pack_layer_hits(rays_o, rays_d, hit, depth):
    """K-layer hits -> uncompacted RaySamplesPacked (slot layout of
    RaySamplerGPU.cuh:206-271 with M = K) -> caller compacts.

    hit [N,K] bool, depth [N,K] f32, both in mesh order (0 = innermost,
    volsurfs.py:476-518).  Slot j of ray r holds that ray's j-th hit in
    outer -> inner order (descending mesh index, volsurfs.py:601-603).
    samples_z = t, samples_3d = the reference kernel's ``positions`` (bvh.cu:445:
    ``ray_o + depth*ray_d``, which nvcc contracts to one fp32 FMA per component —
    see raytrace_oracle.c, contract "device"), samples_dirs = d.

    Returns (RaySamplesPackedNP uncompacted, layer_of_slot [N*K] i32 (-1 unused)).
    """
    rays_o = np.asarray(rays_o, np.float32)
    rays_d = np.asarray(rays_d, np.float32)
    hit = np.asarray(hit, bool)
    depth = np.asarray(depth, np.float32)
    N, K = hit.shape
    rsp = RaySamplesPackedNP(N, N * K, 0, 1)
    rsp.is_compacted = False
    rsp.ray_o = rays_o.copy()
    rsp.ray_d = rays_d.copy()
    layer_of_slot = np.full(N * K, -1, np.int32)
    hf = hit[:, ::-1]
    cnt = hf.sum(axis=1)
    ray, j = np.nonzero(hf)
    layer = K - 1 - j
    rank = (np.cumsum(hf, axis=1) - 1)[ray, j]
    slot = ray * K + rank
    t = depth[ray, layer]
    rsp.samples_z[slot, 0] = t
    from .raytrace import fmaf

    rsp.samples_3d[slot] = fmaf(rays_d[ray], t[:, None], rays_o[ray])
    rsp.samples_dirs[slot] = rays_d[ray]
    layer_of_slot[slot] = layer
    has = cnt > 0
    rsp.ray_start_end_idx[has, 0] = (np.nonzero(has)[0] * K).astype(np.int32)
    rsp.ray_start_end_idx[has, 1] = (np.nonzero(has)[0] * K + cnt[has]).astype(np.int32)
    return rsp, layer_of_slot
