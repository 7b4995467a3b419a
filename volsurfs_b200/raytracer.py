"""K-layer shell intersection + hit packing (host side).

``pack_layer_hits`` turns the layer-major hit records of a K-layer trace into a compacted
:class:`~volsurfs_b200.volsurfs.RaySamplesPacked` (outer -> inner order per ray) — the packed counterpart of the
dense bookkeeping at volsurfs_py/methods/volsurfs.py:476-518,601-603.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .volsurfs import RaySamplesPacked, _stream


def pack_layer_hits(rays_o, rays_d, depth, tri=None, bary_u=None, bary_v=None, t_far: float = 100.0, exact_size: bool = True):
    """rays_o, rays_d [N,3] f32; depth [K,N] f32 (mesh 0 = innermost; miss = 1e6); tri [K,N] i32; bary_u/v [K,N] f32.

    A layer is hit iff ``depth <= t_far`` (raytracelib/raytracer.py:100).  Returns a compacted RaySamplesPacked whose
    ``samples_idx`` is the source slot ``r*K + rank`` and which carries ``samples_layer``, ``samples_triangle`` and
    ``samples_uv`` for the appearance stage.  ``exact_size=True`` sizes the sample arrays to the true total (one
    device->host read, like the reference's compaction); ``False`` keeps capacity N*K arrays and no sync (the arrays'
    tails are unused; ``rsp.total_dev`` holds the count on the device)."""
    L = _lib.lib()
    K, n_rays = int(depth.shape[0]), int(depth.shape[1])
    dev = depth.device
    st = _stream()
    rays_o = rays_o.contiguous()
    rays_d = rays_d.contiguous()
    depth = depth.contiguous()
    scratch = torch.empty(max(int(L.vs_pack_scratch_bytes(n_rays)), 8), dtype=torch.uint8, device=dev)
    total_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    check(L.vs_pack_hits_offsets(ptr(depth), K, float(t_far), n_rays, ptr(total_dev), ptr(scratch), st), "vs_pack_hits_offsets")
    cap = int(total_dev.item()) if exact_size else n_rays * K
    f = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    out = RaySamplesPacked._from_tensors(
        samples_idx=torch.empty((cap, 1), **i32),
        samples_3d=torch.empty((cap, 3), **f),
        samples_dirs=torch.empty((cap, 3), **f),
        samples_z=torch.empty((cap, 1), **f),
        samples_dt=torch.full((cap, 1), -1.0, **f),
        samples_values=torch.full((cap, 1), -1.0, **f),
        ray_start_end_idx=torch.empty((n_rays, 2), **i32),
        ray_o=rays_o,
        ray_d=rays_d,
        ray_enter=torch.full((n_rays, 1), -1.0, **f),
        ray_exit=torch.full((n_rays, 1), -1.0, **f),
        ray_max_dt=torch.full((n_rays, 1), -1.0, **f),
    )
    out.samples_layer = torch.empty((cap,), **i32)
    out.samples_triangle = torch.empty((cap,), **i32) if tri is not None else None
    out.samples_uv = torch.empty((cap, 2), **f) if bary_u is not None and bary_v is not None else None
    out.total_dev = total_dev
    if n_rays > 0:
        check(
            L.vs_pack_hits_scatter(
                ptr(rays_o), ptr(rays_d), ptr(depth), ptr(None if tri is None else tri.contiguous()),
                ptr(None if bary_u is None else bary_u.contiguous()), ptr(None if bary_v is None else bary_v.contiguous()),
                ptr(scratch), K, float(t_far), ptr(out.ray_start_end_idx), ptr(out.samples_idx), ptr(out.samples_3d),
                ptr(out.samples_dirs), ptr(out.samples_z), ptr(out.samples_layer), ptr(out.samples_triangle), ptr(out.samples_uv),
                n_rays, st,
            ),
            "vs_pack_hits_scatter",
        )
    return out
