"""ORACLE (test infrastructure only): ctypes wrapper around raytrace_oracle.c — the CPU restatement of
raytracelib's ``RayTracer`` (raytracelib/raytracer.py:7-113) for K nested meshes."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(_build.build()))
        lib.vso_bvh_build.restype = ctypes.c_void_p
        lib.vso_bvh_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.vso_bvh_free.argtypes = [ctypes.c_void_p]
        lib.vso_bvh_num_nodes.restype = ctypes.c_int
        lib.vso_bvh_num_nodes.argtypes = [ctypes.c_void_p]
        lib.vso_trace.restype = ctypes.c_int
        lib.vso_trace.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int64] + [ctypes.c_void_p] * 8
        lib.vso_set_contract.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def fmaf(a, b, c):
    """element-wise float32 fma(a, b, c) (one rounding) through the C oracle library; arrays are broadcast first"""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float32), np.asarray(b, np.float32), np.asarray(c, np.float32))
    a, b, c = (np.ascontiguousarray(x) for x in (a, b, c))
    out = np.empty(a.shape, np.float32)
    lib = _load()
    lib.vso_fmaf_array.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64]
    lib.vso_fmaf_array(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, a.size)
    return out


class OracleRayTracer:
    """Same call shape as raytracelib.RayTracer: one BVH per mesh, ``trace(rays_o, rays_d, mesh_id)`` returns the
    reference's result dict (numpy).  ``mode='bvh'`` follows the reference traversal, ``mode='brute'`` tests every
    triangle in index order.

    ``contract``: ``"device"`` (default) = the FMA contractions of the reference's CUDA kernel (what the product is held to, bit for
    bit); ``"host"`` = no contraction = the reference's ``__host__`` code path under gcc.  Both are pinned by
    oracle/_ref/libraytrace_ref.so (the reference's own src/bvh.cu), see raytrace_oracle.c."""

    def __init__(self, meshes, t_near=1e-3, t_far=100.0, contract="device"):
        lib = _load()
        assert contract in ("device", "host")
        self.contract = 1 if contract == "device" else 0
        self.t_near, self.t_far = t_near, t_far
        self.handles = []
        self.nr_meshes = len(meshes)
        for verts, faces in meshes:
            v = np.ascontiguousarray(verts, np.float32)
            f = np.ascontiguousarray(faces, np.int32)
            assert f.shape[0] > 8, "BVH needs at least 8 triangles."  # raytracer.py:17
            self.handles.append(lib.vso_bvh_build(v.ctypes.data, f.ctypes.data, f.shape[0]))

    def __del__(self):
        try:
            for h in self.handles:
                _load().vso_bvh_free(h)
        except Exception:  # noqa: BLE001
            pass

    def num_nodes(self, mesh_id=0):
        return _load().vso_bvh_num_nodes(self.handles[mesh_id])

    def trace(self, rays_o, rays_d, mesh_id=0, mode="bvh", min_depth=None):
        o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        md = np.zeros(n, np.float32) if min_depth is None else np.ascontiguousarray(min_depth, np.float32)  # raytracer.py:70
        positions = np.zeros((n, 3), np.float32)
        normals = np.zeros((n, 3), np.float32)
        depth = np.zeros(n, np.float32)
        tmid = np.zeros(n, np.int64)
        tid = np.zeros(n, np.int64)
        bary = np.zeros((n, 3), np.float32)
        u = np.zeros(n, np.float32)
        v = np.zeros(n, np.float32)
        _load().vso_set_contract(self.contract)
        ov = _load().vso_trace(self.handles[mesh_id], 1 if mode == "brute" else 0, o.ctypes.data, d.ctypes.data, md.ctypes.data, n,
                               positions.ctypes.data, normals.ctypes.data, depth.ctypes.data, tmid.ctypes.data, tid.ctypes.data,
                               bary.ctypes.data, u.ctypes.data, v.ctypes.data)
        is_hit = depth <= self.t_far  # raytracer.py:100
        return {"any_hit": bool(is_hit.any()), "is_hit": is_hit, "positions": positions, "triangles_mesh_id": tmid, "triangles_id": tid,
                "depth": depth, "normals": normals, "barycentric": bary, "view_dirs": d, "u": u, "v": v, "stack_overflow": bool(ov)}

    def trace_layers(self, rays_o, rays_d, mode="bvh"):
        """all K meshes: layer-major arrays [K,N] as the CUDA intersector returns them"""
        res = [self.trace(rays_o, rays_d, k, mode) for k in range(self.nr_meshes)]
        return {
            "depth": np.stack([r["depth"] for r in res]),
            "tri": np.stack([r["triangles_id"] for r in res]).astype(np.int32),
            "u": np.stack([r["u"] for r in res]),
            "v": np.stack([r["v"] for r in res]),
            "is_hit": np.stack([r["is_hit"] for r in res]),
            "per_mesh": res,
        }
