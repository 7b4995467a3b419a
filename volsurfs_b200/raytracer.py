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
    # the reference fills unused fields with -1 (RaySamplesPacked.cu:13-48); the capacity-mode fast path leaves them uninitialised
    filled = (lambda shape: torch.full(shape, -1.0, **f)) if exact_size else (lambda shape: torch.empty(shape, **f))
    out = RaySamplesPacked._from_tensors(
        samples_idx=torch.empty((cap, 1), **i32),
        samples_3d=torch.empty((cap, 3), **f),
        samples_dirs=torch.empty((cap, 3), **f),
        samples_z=torch.empty((cap, 1), **f),
        samples_dt=filled((cap, 1)),
        samples_values=filled((cap, 1)),
        ray_start_end_idx=torch.empty((n_rays, 2), **i32),
        ray_o=rays_o,
        ray_d=rays_d,
        ray_enter=filled((n_rays, 1)),
        ray_exit=filled((n_rays, 1)),
        ray_max_dt=filled((n_rays, 1)),
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


class ShellTracer:
    """K nested meshes on the GPU: the B200 counterpart of ``raytracelib.RayTracer``
    (submodules/raytracelib/raytracelib/raytracer.py:7-223).

    ``meshes``: list of objects with ``.vertices`` / ``.faces`` (as raytracelib expects) or ``(vertices, faces)`` tuples,
    mesh 0 = innermost (mesh file order, volsurfs_py/utils/mesh_loaders.py:22-31).  The BVHs are built once on the host.

    * :meth:`trace_layers` — all K layers in ONE launch, layer-major compact records (depth, face, u, v);
    * :meth:`render_samples` — trace + pack into a compacted :class:`RaySamplesPacked` (outer -> inner per ray);
    * :meth:`trace` — reference-compatible per-mesh result dict (same keys, dtypes and shapes as raytracer.py:103-113),
      without the reference's ``torch.cuda.synchronize()``.
    """

    def __init__(self, meshes, t_near: float = 1e-3, t_far: float = 100.0):
        import ctypes

        import numpy as np

        L = _lib.lib()
        self.t_near, self.t_far = t_near, t_far
        self.nr_meshes = len(meshes)
        vs, fs = [], []
        for m in meshes:
            v, f = (m.vertices, m.faces) if hasattr(m, "vertices") else m
            if torch.is_tensor(v):
                v = v.detach().cpu().numpy()
            if torch.is_tensor(f):
                f = f.detach().cpu().numpy()
            v = np.ascontiguousarray(v, dtype=np.float32)
            f = np.ascontiguousarray(f, dtype=np.int32)
            assert f.shape[0] > 8, "BVH needs at least 8 triangles."  # raytracer.py:17
            vs.append(v)
            fs.append(f)
        K = self.nr_meshes
        vptr = (ctypes.c_void_p * K)(*[v.ctypes.data for v in vs])
        fptr = (ctypes.c_void_p * K)(*[f.ctypes.data for f in fs])
        nv = (ctypes.c_int64 * K)(*[v.shape[0] for v in vs])
        nf = (ctypes.c_int64 * K)(*[f.shape[0] for f in fs])
        handle = ctypes.c_void_p()
        if not torch.cuda.is_available():
            raise _lib.VolsurfsB200Error("ShellTracer needs a CUDA device (no CPU fallback)")
        check(L.vs_shells_build(K, vptr, nv, fptr, nf, ctypes.byref(handle)), "vs_shells_build")
        self._handle = handle
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_faces = [int(f.shape[0]) for f in fs]

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _lib.lib().vs_shells_free(h)
            except Exception:  # noqa: BLE001
                pass
            self._handle = None

    def num_nodes(self, layer: int = 0) -> int:
        import ctypes

        n = ctypes.c_int64()
        check(_lib.lib().vs_shells_info(self._handle, layer, ctypes.byref(n), None), "vs_shells_info")
        return int(n.value)

    def overflowed(self) -> bool:
        return bool(_lib.lib().vs_shells_overflowed(self._handle))

    @staticmethod
    def _rays(rays_o, rays_d):
        assert torch.is_tensor(rays_o) and rays_o.is_cuda, "rays_o must be a torch.Tensor on cuda"
        assert torch.is_tensor(rays_d) and rays_d.is_cuda, "rays_d must be a torch.Tensor on cuda"
        return rays_o.float().contiguous().view(-1, 3), rays_d.float().contiguous().view(-1, 3)

    @torch.no_grad()
    def trace_layers(self, rays_o, rays_d, layer_first: int = 0, layer_count=None):
        """-> dict(depth [k,N] f32 (1e6 = miss), tri [k,N] i32 (-1 = miss), u, v [k,N] f32) for layers
        [layer_first, layer_first + layer_count)"""
        rays_o, rays_d = self._rays(rays_o, rays_d)
        k = self.nr_meshes - layer_first if layer_count is None else layer_count
        n = rays_o.shape[0]
        dev = rays_o.device
        depth = torch.empty((k, n), dtype=torch.float32, device=dev)
        tri = torch.empty((k, n), dtype=torch.int32, device=dev)
        u = torch.empty((k, n), dtype=torch.float32, device=dev)
        v = torch.empty((k, n), dtype=torch.float32, device=dev)
        check(
            _lib.lib().vs_shells_trace(self._handle, ptr(rays_o), ptr(rays_d), n, layer_first, k, ptr(depth), ptr(tri), ptr(u), ptr(v), _stream()),
            "vs_shells_trace",
        )
        return {"depth": depth, "tri": tri, "u": u, "v": v, "rays_o": rays_o, "rays_d": rays_d}

    @torch.no_grad()
    def render_samples(self, rays_o, rays_d, exact_size: bool = True, with_normals: bool = True):
        """trace all layers and pack the hits: compacted RaySamplesPacked (+ ``samples_normals`` [S,3])"""
        rec = self.trace_layers(rays_o, rays_d)
        rsp = pack_layer_hits(rec["rays_o"], rec["rays_d"], rec["depth"], rec["tri"], rec["u"], rec["v"], t_far=self.t_far,
                              exact_size=exact_size)
        if with_normals:
            S = rsp.get_max_nr_samples()
            alloc = torch.zeros if exact_size else torch.empty
            rsp.samples_normals = alloc((S, 3), dtype=torch.float32, device=rec["depth"].device)
            check(
                _lib.lib().vs_shells_sample_normals(self._handle, ptr(rsp.samples_layer), ptr(rsp.samples_triangle), S,
                                                    None if exact_size else ptr(rsp.total_dev), ptr(rsp.samples_normals), _stream()),
                "vs_shells_sample_normals",
            )
        return rsp

    @torch.no_grad()
    def set_face_uvs(self, face_uvs_per_layer):
        """per-face-vertex texture coordinates of every layer ([F_l,3,2] each; ``TensorMesh.get_faces_uvs()`` in the reference)"""
        dev = self.device
        tabs = [torch.as_tensor(t, dtype=torch.float32).reshape(-1, 3, 2) for t in face_uvs_per_layer]
        assert len(tabs) == self.nr_meshes
        offs, o = [], 0
        for t in tabs:
            offs.append(o)
            o += int(t.shape[0])
        self._face_uvs = torch.cat(tabs).contiguous().to(dev)
        self._face_uv_offset = torch.tensor(offs, dtype=torch.int32, device=dev)

    @torch.no_grad()
    def sample_uvs(self, rsp):
        """texture coordinates of the packed hits, uv = sum_j barycentric_j * face_uv_j (volsurfs.py:509-516) -> [S,2]"""
        assert getattr(self, "_face_uvs", None) is not None, "call set_face_uvs first"
        S = rsp.get_max_nr_samples()
        out = torch.zeros((S, 2), dtype=torch.float32, device=self.device)
        check(_lib.lib().vs_shells_sample_uvs(ptr(rsp.samples_layer), ptr(rsp.samples_triangle), ptr(rsp.samples_uv), ptr(self._face_uvs),
                                              ptr(self._face_uv_offset), S, ptr(getattr(rsp, "total_dev", None)), ptr(out), _stream()),
              "vs_shells_sample_uvs")
        return out

    @torch.no_grad()
    def trace(self, rays_o, rays_d, mesh_id: int = 0, **_unused):
        """Reference-compatible single-mesh trace (raytracer.py:35-113): same dict keys / dtypes / shapes."""
        assert mesh_id < self.nr_meshes, "mesh_id must be smaller than the number of meshes in the scene"
        rec = self.trace_layers(rays_o, rays_d, layer_first=mesh_id, layer_count=1)
        rays_o, rays_d = rec["rays_o"], rec["rays_d"]
        n = rays_o.shape[0]
        dev = rays_o.device
        positions = torch.empty((n, 3), dtype=torch.float32, device=dev)
        normals = torch.empty((n, 3), dtype=torch.float32, device=dev)
        bary = torch.empty((n, 3), dtype=torch.float32, device=dev)
        tmid = torch.empty((n,), dtype=torch.int64, device=dev)
        tid = torch.empty((n,), dtype=torch.int64, device=dev)
        depth = rec["depth"][0]
        check(
            _lib.lib().vs_shells_expand(self._handle, mesh_id, ptr(rays_o), ptr(rays_d), ptr(depth), ptr(rec["tri"][0]), ptr(rec["u"][0]),
                                        ptr(rec["v"][0]), n, ptr(positions), ptr(normals), ptr(tmid), ptr(tid), ptr(bary), _stream()),
            "vs_shells_expand",
        )
        is_hit = depth <= self.t_far          # raytracer.py:100
        return {
            "any_hit": is_hit.any(dim=-1), "is_hit": is_hit, "positions": positions, "triangles_mesh_id": tmid, "triangles_id": tid,
            "depth": depth, "normals": normals, "barycentric": bary, "view_dirs": rays_d,
        }

    def trace_all(self, rays_o, rays_d, **kw):
        return [self.trace(rays_o, rays_d, mesh_id=i) for i in range(self.nr_meshes)]
