"""ORACLE (test infrastructure only): ctypes wrapper around oracle/_ref/libraytrace_ref.so — the REFERENCE's own mesh ray tracer
(submodules/raytracelib/src/bvh.cu, compiled where it lies by oracle/build.py:build_ref_raytrace) with the call shape of
raytracelib.RayTracer (raytracelib/raytracer.py:7-113).  ``trace_host`` runs the reference's ``__host__ __device__`` traversal on the
CPU (numpy in/out); ``trace_gpu`` launches the reference's CUDA kernel on torch CUDA tensors."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build

_lib = None


def available() -> bool:
    return _build.raytrace_ref_available()


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(str(_build.RAYTRACE_LIB))
        lib.ref_rt_create.restype = ctypes.c_void_p
        lib.ref_rt_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        lib.ref_rt_free.argtypes = [ctypes.c_void_p]
        lib.ref_rt_num_nodes.argtypes = [ctypes.c_void_p]
        lib.ref_rt_trace.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int]
        lib.ref_rt_trace_host.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int]
        _lib = lib
    return _lib


class RefRayTracer:
    def __init__(self, meshes, t_far=100.0, gpu=False):
        lib = _load()
        self.t_far = t_far
        self.gpu = gpu
        self.handles = []
        for verts, faces in meshes:
            v = np.ascontiguousarray(verts, np.float32)
            f = np.ascontiguousarray(faces, np.uint32)
            assert f.shape[0] > 8, "BVH needs at least 8 triangles."  # raytracer.py:17
            h = lib.ref_rt_create(v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0], 1 if gpu else 0)
            if not h:
                raise RuntimeError("ref_rt_create failed")
            self.handles.append(h)

    def __del__(self):
        try:
            for h in self.handles:
                _load().ref_rt_free(h)
        except Exception:  # noqa: BLE001
            pass

    def num_nodes(self, mesh_id=0):
        return _load().ref_rt_num_nodes(self.handles[mesh_id])

    def trace_host(self, rays_o, rays_d, mesh_id=0):
        o = np.ascontiguousarray(rays_o, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(rays_d, np.float32).reshape(-1, 3)
        n = o.shape[0]
        md = np.zeros(n, np.float32)  # raytracer.py:70
        out = {"positions": np.zeros((n, 3), np.float32), "normals": np.zeros((n, 3), np.float32), "depth": np.zeros(n, np.float32),
               "triangles_mesh_id": np.zeros(n, np.int64), "triangles_id": np.zeros(n, np.int64), "barycentric": np.zeros((n, 3), np.float32)}
        rc = _load().ref_rt_trace_host(self.handles[mesh_id], o.ctypes.data, d.ctypes.data, md.ctypes.data, out["positions"].ctypes.data,
                                       out["normals"].ctypes.data, out["depth"].ctypes.data, out["triangles_mesh_id"].ctypes.data,
                                       out["triangles_id"].ctypes.data, out["barycentric"].ctypes.data, n)
        assert rc == 0
        out["is_hit"] = out["depth"] <= self.t_far  # raytracer.py:100
        return out

    def trace_gpu(self, rays_o, rays_d, mesh_id=0):
        """rays: torch CUDA float32 [N,3] contiguous; returns torch tensors (the reference's pre-allocated outputs, raytracer.py:72-95)"""
        import torch

        assert self.gpu and rays_o.is_cuda and rays_o.is_contiguous() and rays_d.is_contiguous()
        n = rays_o.shape[0]
        dev = rays_o.device
        md = torch.zeros(n, dtype=torch.float32, device=dev)
        out = {"positions": torch.zeros(n, 3, device=dev), "normals": torch.zeros(n, 3, device=dev), "depth": torch.zeros(n, device=dev),
               "triangles_mesh_id": torch.zeros(n, dtype=torch.int64, device=dev), "triangles_id": torch.zeros(n, dtype=torch.int64, device=dev),
               "barycentric": torch.zeros(n, 3, device=dev)}
        torch.cuda.synchronize()
        rc = _load().ref_rt_trace(self.handles[mesh_id], rays_o.data_ptr(), rays_d.data_ptr(), md.data_ptr(), out["positions"].data_ptr(),
                                  out["normals"].data_ptr(), out["depth"].data_ptr(), out["triangles_mesh_id"].data_ptr(),
                                  out["triangles_id"].data_ptr(), out["barycentric"].data_ptr(), n)
        if rc != 0:
            raise RuntimeError(f"reference raytrace kernel failed: cuda error {rc}")
        out["is_hit"] = out["depth"] <= self.t_far
        return out
