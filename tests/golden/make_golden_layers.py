"""Golden vectors for the K-layer hit bookkeeping, produced by EXECUTING THE REFERENCE'S OWN SOURCE LINES
(volsurfs_py/methods/volsurfs.py:449-488 result buffers + trace loop, :492-516 per-mesh scatter of hits / points / normals / uvs) on CPU
tensors.  Run where /root/reference is mounted; the .npz is committed, the reference source is not.

    python tests/golden/make_golden_layers.py

`self.raytracer.trace` is served by the mesh-tracer oracle in the reference kernel's arithmetic (oracle/raytrace_oracle.c, contract
"device": bit-identical to the reference's CUDA kernel, tests/test_gpu_raytrace.py) returning the result dict of
raytracelib/raytracer.py:103-113 as torch tensors; `self.tensor_meshes[i].get_faces_uvs()` returns the synthetic shells' [F,3,2] chart.
layers_k3.npz: rays, and the reference's dense buffers surfs_hits [N,K], surfs_points [N,K,3], surfs_normals [N,K,3], surfs_uvs [N,K,2].
"""
from __future__ import annotations

import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))

from oracle.raytrace import OracleRayTracer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays, shell_face_uvs, shell_meshes  # noqa: E402

K, N_LAT, N_LON, RES = 3, 48, 48, 64


def block(first: int, last: int) -> str:
    lines = (REF / "volsurfs_py/methods/volsurfs.py").read_text().splitlines()[first - 1:last]
    return textwrap.dedent("\n".join(lines))


class _Tracer:
    """raytracelib.RayTracer.trace call shape (raytracer.py:35-113) over the oracle"""

    def __init__(self, meshes):
        self.oracle = OracleRayTracer(meshes, contract="device")

    def trace(self, rays_o, rays_d, mesh_id=0):
        r = self.oracle.trace(rays_o.numpy(), rays_d.numpy(), mesh_id)
        out = {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("is_hit", "positions", "triangles_mesh_id", "triangles_id", "depth",
                                                                        "normals", "barycentric")}
        out["any_hit"] = torch.tensor(r["any_hit"])
        out["view_dirs"] = rays_d
        return out


def main():
    meshes = shell_meshes(K=K, n_lat=N_LAT, n_lon=N_LON)
    rays_o, rays_d = camera_rays(RES, RES)
    face_uvs = torch.from_numpy(shell_face_uvs(N_LAT, N_LON))
    tm = types.SimpleNamespace(get_faces_uvs=lambda: face_uvs)
    self_ = types.SimpleNamespace(nr_meshes=K, profiler=None, raytracer=_Tracer(meshes), tensor_meshes=[tm] * K,
                                  hyper_params=types.SimpleNamespace(using_neural_textures=True))
    ns = {"torch": torch, "self": self_, "rays_o": rays_o, "rays_d": rays_d, "nr_rays": rays_o.shape[0], "debug_ray_idx": None}
    exec(block(449, 488), ns)   # buffers + `for i in range(self.nr_meshes): res = self.raytracer.trace(...)`
    # volsurfs.py:492-516: the head of the shading loop's body (the loop statement itself is :492-493), up to the uv scatter
    exec(block(492, 516), ns)
    out = {"rays_o": rays_o.numpy(), "rays_d": rays_d.numpy(), "surfs_hits": ns["surfs_hits"].numpy(), "surfs_points": ns["surfs_points"].numpy(),
           "surfs_normals": ns["surfs_normals"].numpy(), "surfs_uvs": ns["surfs_uvs"].numpy(),
           "params": np.array([K, N_LAT, N_LON, RES])}
    np.savez_compressed(OUT / "layers_k3.npz", **out)
    print("hits per layer:", out["surfs_hits"].sum(0), "of", rays_o.shape[0])


if __name__ == "__main__":
    main()
