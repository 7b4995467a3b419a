"""Shared synthetic inputs of the sampler tests: camera-like rays against a cuboid grid whose occupancy is a lumpy blob inside a
spherical region of interest (a NeRF-synthetic-like occupied core surrounded by empty space)."""
import numpy as np

from oracle import sampler as osamp


def make_scene(n_rays=300, n=16, extent=(1.0, 1.2, 0.9), seed=0):
    rs = np.random.RandomState(seed)
    extent = np.asarray(extent, np.float32)
    # ray origins on a sphere of radius 2, directions towards jittered points near the origin (some rays miss the box)
    o = rs.randn(n_rays, 3)
    o = (2.0 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    target = (rs.rand(n_rays, 3) - 0.5) * 1.4 * extent
    d = target - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[: min(3, n_rays)] = np.array([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], np.float32)[: min(3, n_rays)]
    o[: min(3, n_rays)] = np.array([[0.01, 0.02, 2.0], [-2.0, 0.03, 0.01], [0.02, 2.0, -0.03]], np.float32)[: min(3, n_rays)]
    # slab test against the cuboid, shrunk so that every marched position stays inside the grid
    half = extent / 2 * 0.999
    with np.errstate(divide="ignore", invalid="ignore"):
        t1, t2 = (-half - o) / d, (half - o) / d
    tn = np.nanmax(np.minimum(t1, t2), axis=1)
    tf = np.nanmin(np.maximum(t1, t2), axis=1)
    hit = (tf > tn) & (tf > 0)
    t_entry = np.where(hit, np.maximum(tn, 0) + 1e-4, 0).astype(np.float32).reshape(-1, 1)
    t_exit = np.where(hit, tf - 1e-4, 0).astype(np.float32).reshape(-1, 1)
    # occupancy / roi in Morton order
    vals = rs.rand(n ** 3).astype(np.float32)
    ix, iy, iz = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()

    def spread(v):  # 10-bit Morton spreading (enough for n <= 1024); checked against oracle.sampler.morton3d below
        v = v.astype(np.uint32)
        v = (v | (v << 16)) & np.uint32(0x030000FF)
        v = (v | (v << 8)) & np.uint32(0x0300F00F)
        v = (v | (v << 4)) & np.uint32(0x030C30C3)
        v = (v | (v << 2)) & np.uint32(0x09249249)
        return v

    m = (spread(ix) | (spread(iy) << 1) | (spread(iz) << 2)).astype(np.int64)
    for k in (0, 1, n * n + 3, n ** 3 - 1):
        assert m[k] == osamp.morton3d(int(ix[k]), int(iy[k]), int(iz[k]))
    g = (np.arange(n) + 0.5) / n - 0.5
    p = np.stack([g[ix], g[iy], g[iz]], 1)
    occ = np.zeros(n ** 3, bool)
    roi = np.zeros(n ** 3, bool)
    roi[m] = np.linalg.norm(p, axis=1) < 0.48
    occ[m] = (np.linalg.norm(p * np.array([1.0, 1.3, 0.8]), axis=1) < 0.3 + 0.08 * np.sin(9 * p[:, 0]) * np.cos(7 * p[:, 1])) | (rs.rand(n ** 3) < 0.02)
    return dict(o=o, d=d, t_entry=t_entry, t_exit=t_exit, n=n, extent=extent, occ=occ, roi=roi, vals=vals)
