"""Golden vectors for the baked-texture MeshRenderer, produced by EXECUTING THE REFERENCE'S OWN SOURCE: the bodies of ``render_rays`` and
``shade`` of volsurfs_py/renderers/mesh_renderer.py (:62-110, :112-201) with the reference's own ``TensorTexture``
(submodules/mvdatasets/mvdatasets/utils/tensor_texture.py) and ``SHEncoder.eval`` (volsurfs_py/encodings/sphericalharmonics.py), on CPU
tensors.  Run where /root/reference is mounted; the .npz is committed, the reference sources are not.

    python tests/golden/make_golden_mesh_renderer.py

Stand-ins: ``cv2`` (imported by mvdatasets.utils.images, unused on this path) is an empty module; the literal device string "cuda" in the
two method bodies becomes "cpu"; ``self.raytracer.trace`` is served by the mesh-tracer oracle in the reference kernel's arithmetic
(bit-identical to the reference's CUDA kernel, tests/test_gpu_raytrace.py); ``self.tensor_mesh.get_faces_uvs()`` returns the synthetic
shell's chart.  mesh_renderer_deg{0,3}.npz: mesh parameters, texture, rays and the six shaded buffers.
"""
from __future__ import annotations

import inspect
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

N_LAT, N_LON, RES = 48, 48, 72


def reference_methods():
    import importlib.util

    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    # mvdatasets/__init__.py pulls in dataset loaders (pycolmap, ...): load the files this path uses directly, under their own names
    for pkg in ("mvdatasets", "mvdatasets.utils"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    printing = types.ModuleType("mvdatasets.utils.printing")
    printing.print_error = printing.print_warning = print
    sys.modules["mvdatasets.utils.printing"] = printing
    for name in ("images", "tensor_texture"):
        spec = importlib.util.spec_from_file_location(f"mvdatasets.utils.{name}", REF / f"submodules/mvdatasets/mvdatasets/utils/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"mvdatasets.utils.{name}"] = mod
        spec.loader.exec_module(mod)
    TensorTexture = sys.modules["mvdatasets.utils.tensor_texture"].TensorTexture
    import scipy.special

    if not hasattr(scipy.special, "sph_harm"):  # the reference imports a name newer SciPy removed; SHEncoder.eval never uses it
        scipy.special.sph_harm = scipy.special.sph_harm_y
    sys.modules.setdefault("permutohedral_encoding", types.ModuleType("permutohedral_encoding"))  # imported by the package, unused here
    sys.path.insert(0, str(REF))
    spec = importlib.util.spec_from_file_location("ref_sphericalharmonics", REF / "volsurfs_py/encodings/sphericalharmonics.py")
    shmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(shmod)
    SHEncoder = shmod.SHEncoder

    lines = (REF / "volsurfs_py/renderers/mesh_renderer.py").read_text().splitlines()
    src = textwrap.dedent("\n".join(lines[61:201])).replace('"cuda"', '"cpu"')      # :62-201 = shade + render_rays
    ns = {"torch": torch, "SHEncoder": SHEncoder, "print": lambda *a, **k: None}
    exec(src, ns)
    return TensorTexture, ns["shade"], ns["render_rays"]


class _Tracer:
    def __init__(self, mesh):
        self.oracle = OracleRayTracer([mesh], contract="device")

    def trace(self, rays_o, rays_d):
        r = self.oracle.trace(rays_o.numpy(), rays_d.numpy(), 0)
        out = {k: torch.from_numpy(np.ascontiguousarray(r[k])) for k in ("is_hit", "positions", "triangles_id", "depth", "normals", "barycentric")}
        out["any_hit"] = torch.tensor(r["any_hit"])
        return out


def make(name, nr_coeffs, tex_res, seed):
    TensorTexture, shade, render_rays = reference_methods()
    mesh = shell_meshes(K=1, n_lat=N_LAT, n_lon=N_LON)[0]
    face_uvs = shell_face_uvs(N_LAT, N_LON)
    rng = np.random.default_rng(seed)
    texture = (rng.standard_normal((tex_res[0], tex_res[1], 4 * nr_coeffs)) * 1.5).astype(np.float32)
    o, d = camera_rays(RES, RES)
    self_ = types.SimpleNamespace(raytracer=_Tracer(mesh), tensor_mesh=types.SimpleNamespace(get_faces_uvs=lambda: torch.from_numpy(face_uvs)),
                                  tensor_texture=TensorTexture(texture_np=texture, lerp=True, device="cpu"),
                                  bg_color=torch.tensor((255, 255, 255), dtype=torch.float32) / 255.0)
    self_.shade = types.MethodType(shade.__wrapped__ if hasattr(shade, "__wrapped__") else shade, self_)
    with torch.no_grad():
        res = render_rays.__wrapped__(self_, o, d) if hasattr(render_rays, "__wrapped__") else render_rays(self_, o, d)
    bufs = res["renders"]["ray_traced"]
    np.savez_compressed(OUT / f"mesh_renderer_{name}.npz", texture=texture, rays_o=o.numpy(), rays_d=d.numpy(),
                        params=np.array([N_LAT, N_LON, nr_coeffs]), **{k: v.numpy() for k, v in bufs.items()})
    print(name, {k: tuple(v.shape) for k, v in bufs.items()}, "hits", int(bufs["is_hit"].sum()))


if __name__ == "__main__":
    make("deg3", 16, (40, 56), 1)
    make("deg0", 1, (64, 64), 2)
