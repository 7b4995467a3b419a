"""Golden vectors for the mesh ray tracer, produced by RUNNING THE REFERENCE'S OWN TRACER (submodules/raytracelib/src/bvh.cu compiled
where it lies into oracle/_ref/libraytrace_ref.so, see oracle/build.py:build_ref_raytrace) on the reference's own test meshes.

    python tests/golden/make_golden_raytrace.py            # here (no GPU): meshes + host-path goldens
    python tests/golden/make_golden_raytrace.py --device   # on the GPU box: adds the reference CUDA kernel's outputs
                                                           # (written to gpurun_out/, then committed under tests/golden/)

mesh_{smurf,plushy}.npz
    Geometry of submodules/raytracelib/meshes/{smurf,plushy}.obj — the two meshes of the reference's only tracer test
    (submodules/raytracelib/tests/test_raytracing.py:5) — as float32 vertices / int32 faces (`v` and `f` records; texture indices
    dropped).  The .obj files are test DATA, not source; they stay where they are, only this compact form is committed.
raytrace_{smurf,plushy}_host.npz
    rays (the probe ray of test_raytracing.py:18-19 first, then a 160x160 pinhole view of the mesh) and the outputs of the
    reference's __host__ traversal (TriangleBvh4::ray_intersect + the body of raytrace_kernel) on them: gcc, no FMA contraction.
raytrace_{smurf,plushy}_device.npz
    the same rays through the reference's CUDA kernel (nvcc -O3, sm_100a): what the product kernel is held to bit for bit.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
ROOT = OUT.parent.parent
sys.path.insert(0, str(ROOT))
MESH_DIR = Path("/root/reference/submodules/raytracelib/meshes")
NAMES = ("smurf", "plushy")


def load_obj(path: Path):
    verts, faces = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                faces.append([int(tok.split("/")[0]) - 1 for tok in line.split()[1:4]])
    return np.asarray(verts, np.float64).astype(np.float32), np.asarray(faces, np.int32)


def view_rays(verts: np.ndarray, res: int = 160):
    """probe ray of the reference's test + a pinhole view of the mesh from (+x, slightly up), looking at its centre"""
    lo, hi = verts.min(0), verts.max(0)
    centre = ((lo + hi) / 2).astype(np.float64)
    radius = float(np.linalg.norm(hi - lo)) / 2
    eye = centre + np.array([2.2, 0.7, 0.9]) * radius
    fwd = (centre - eye) / np.linalg.norm(centre - eye)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    s = np.tan(np.deg2rad(15.0))
    px = (np.arange(res) + 0.5) / res * 2 - 1
    gx, gy = np.meshgrid(px, -px)
    d = fwd[None, None] + s * gx[..., None] * right[None, None] + s * gy[..., None] * up[None, None]
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d.reshape(-1, 3)
    o = np.broadcast_to(eye, d.shape)
    o = np.concatenate([[[2.0, 0.0, 0.0]], o]).astype(np.float32)      # tests/test_raytracing.py:18
    d = np.concatenate([[[-1.0, 0.0, 0.0]], d]).astype(np.float32)     # tests/test_raytracing.py:19
    return np.ascontiguousarray(o), np.ascontiguousarray(d)


KEYS = ("depth", "triangles_id", "triangles_mesh_id", "positions", "normals", "barycentric")


def main():
    device = "--device" in sys.argv
    from oracle.ref_raytrace import RefRayTracer

    for name in NAMES:
        mesh_file = OUT / f"mesh_{name}.npz"
        if not mesh_file.exists():
            v, f = load_obj(MESH_DIR / f"{name}.obj")
            np.savez_compressed(mesh_file, verts=v, faces=f)
        m = np.load(mesh_file)
        v, f = m["verts"], m["faces"]
        o, d = view_rays(v)
        if not device:
            rt = RefRayTracer([(v, f)])
            r = rt.trace_host(o, d)
            np.savez_compressed(OUT / f"raytrace_{name}_host.npz", rays_o=o, rays_d=d, nodes=rt.num_nodes(), **{k: r[k] for k in KEYS})
            print(name, "host: verts", v.shape, "faces", f.shape, "nodes", rt.num_nodes(), "hits", int(r["is_hit"].sum()), "of", len(o),
                  "probe ray:", float(r["depth"][0]), int(r["triangles_id"][0]))
        else:
            import torch

            rt = RefRayTracer([(v, f)], gpu=True)
            r = rt.trace_gpu(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
            out_dir = ROOT / "gpurun_out"
            out_dir.mkdir(exist_ok=True)
            np.savez_compressed(out_dir / f"raytrace_{name}_device.npz", rays_o=o, rays_d=d, **{k: r[k].cpu().numpy() for k in KEYS})
            print(name, "device: hits", int(r["is_hit"].sum()), "of", len(o), "probe ray:", float(r["depth"][0]), int(r["triangles_id"][0]))


if __name__ == "__main__":
    main()
