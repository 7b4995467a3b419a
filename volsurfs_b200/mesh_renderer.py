"""Baked-texture mesh renderer: the B200 counterpart of ``volsurfs_py/renderers/mesh_renderer.py:MeshRenderer`` (the real-time viewer path of
a baked volsurfs scene: ONE textured mesh, SH coefficients stored in an RGBA-per-coefficient texture).

``render_rays(rays_o, rays_d)`` returns the reference's ``{"renders": {"ray_traced": {...}}}`` dict (mesh_renderer.py:112-201): the mesh trace
runs on :class:`~volsurfs_b200.raytracer.ShellTracer` (one layer), everything after it — texture coordinates, bilinear lookup of the
zero-padded texture (``TensorTexture(lerp=True)``, mvdatasets/utils/tensor_texture.py), fp16 SH evaluation, sigmoid, and ``shade`` — is one
kernel (``vs_baked_texture_shade``).  Scene loading (scene.json / .obj / texture files through open3d and PIL, mesh_renderer.py:24-45) is
outside the hot path: construct from arrays."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .raytracer import ShellTracer
from .volsurfs import _stream


class MeshRenderer:
    def __init__(self, vertices, faces, face_uvs, texture, t_near: float = 1e-3, t_far: float = 100.0, bg_color=(255, 255, 255), device=None):
        """vertices [V,3] f32, faces [F,3] int, face_uvs [F,3,2] f32 (``TensorMesh.get_faces_uvs()``), texture [H,W,4*nr_coeffs] f32 (the baked
        SH coefficients, ``mesh.texture.image``); bg_color: 8-bit rgb (mesh_renderer.py:55-60)"""
        self.tracer = ShellTracer([(np.asarray(vertices, np.float32), np.asarray(faces, np.int32))], t_near=t_near, t_far=t_far)
        dev = self.tracer.device if device is None else device
        tex = torch.as_tensor(np.asarray(texture), dtype=torch.float32)
        tex = tex.reshape(tex.shape[0], tex.shape[1], -1)
        self.res = (int(tex.shape[0]), int(tex.shape[1]))          # height, width
        assert tex.shape[2] % 4 == 0 and tex.shape[2] // 4 in (1, 4, 9, 16), "texture channels must be 4 * (sh_deg + 1)^2"
        self.nr_coeffs = int(tex.shape[2]) // 4
        padded = torch.zeros((tex.shape[0] + 2, tex.shape[1] + 2, tex.shape[2]), dtype=torch.float32)   # tensor_texture.py:55-64
        padded[1:-1, 1:-1] = tex
        self.texture = padded.to(dev).contiguous()
        self.face_uvs = torch.as_tensor(np.asarray(face_uvs), dtype=torch.float32).reshape(-1, 3, 2).to(dev).contiguous()
        assert self.face_uvs.shape[0] == np.asarray(faces).shape[0]
        self.bg_color = [float(c) / 255.0 for c in bg_color]
        self._bg_c = (_lib.ctypes.c_float * 3)(*self.bg_color)
        self.active_render_mode, self.active_shader = "ray_traced", "rgb"

    @torch.no_grad()
    def render_rays(self, rays_o, rays_d, verbose: bool = False) -> dict:
        res = self.tracer.trace(rays_o, rays_d, mesh_id=0)
        n = int(rays_o.shape[0])
        dev = rays_o.device
        f = dict(dtype=torch.float32, device=dev)
        out = {"is_hit": torch.empty((n, 1), **f), "normals": torch.empty((n, 3), **f), "uvs": torch.empty((n, 3), **f),
               "rgb": torch.empty((n, 3), **f), "alpha": torch.empty((n, 1), **f), "view_dirs": torch.empty((n, 3), **f)}
        hit_u8 = res["is_hit"].to(torch.uint8).contiguous()
        check(_lib.lib().vs_baked_texture_shade(
            ptr(hit_u8), ptr(res["triangles_id"].contiguous()), ptr(res["barycentric"].contiguous()), ptr(rays_d.contiguous()),
            ptr(res["normals"].contiguous()), ptr(self.face_uvs), ptr(self.texture), self.res[0], self.res[1], self.nr_coeffs, self._bg_c,
            ptr(out["is_hit"]), ptr(out["normals"]), ptr(out["uvs"]), ptr(out["rgb"]), ptr(out["alpha"]), ptr(out["view_dirs"]), n,
            _stream()), "vs_baked_texture_shade")
        return {"renders": {"ray_traced": out}}
