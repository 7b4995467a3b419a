"""Baked-texture MeshRenderer (second half) against goldens recorded by executing the reference's own
``MeshRenderer.render_rays`` / ``shade`` source with the reference's ``TensorTexture`` and ``SHEncoder.eval`` on CPU tensors
(tests/golden/make_golden_mesh_renderer.py).  Bars: hit mask, normals, view directions and texture coordinates bit-exact; colour and alpha
within 1e-6 on >= 99.8 % of the hit pixels and within 4e-3 everywhere — the coefficients pass through ``.half()`` after the bilinear sum, so a
last-bit difference of torch's CPU reduction order flips an fp16 rounding (one ulp = 1e-3 of a coefficient of magnitude ~1.5)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["deg3", "deg0"])
def test_mesh_renderer_vs_reference_source_golden(name):
    from volsurfs_b200.mesh_renderer import MeshRenderer
    from volsurfs_b200.synthetic import shell_face_uvs, shell_meshes

    g = np.load(GOLDEN / f"mesh_renderer_{name}.npz")
    n_lat, n_lon, nr_coeffs = (int(x) for x in g["params"])
    verts, faces = shell_meshes(K=1, n_lat=n_lat, n_lon=n_lon)[0]
    r = MeshRenderer(verts, faces, shell_face_uvs(n_lat, n_lon), g["texture"])
    assert r.nr_coeffs == nr_coeffs
    out = r.render_rays(torch.from_numpy(g["rays_o"]).cuda(), torch.from_numpy(g["rays_d"]).cuda())["renders"]["ray_traced"]
    assert set(out) == {"is_hit", "normals", "uvs", "rgb", "alpha", "view_dirs"}
    hit = g["is_hit"][:, 0] > 0
    assert hit.sum() > 1000
    for key in ("is_hit", "normals", "view_dirs", "uvs"):
        assert np.array_equal(out[key].cpu().numpy(), g[key]), key
    for key in ("rgb", "alpha"):
        a, b = out[key].cpu().numpy(), g[key]
        assert a.shape == b.shape
        assert np.array_equal(a[~hit], b[~hit])                         # background colour / zero alpha
        err = np.abs(a[hit] - b[hit]).max(axis=1)
        print(name, key, "max err", float(err.max()), "exact share", float((err <= 1e-6).mean()))
        assert (err <= 1e-6).mean() >= 0.998 and err.max() <= 4e-3


def test_mesh_renderer_all_missed_and_argument_errors():
    from volsurfs_b200.mesh_renderer import MeshRenderer
    from volsurfs_b200.synthetic import shell_face_uvs, shell_meshes

    verts, faces = shell_meshes(K=1, n_lat=16, n_lon=16)[0]
    tex = np.zeros((8, 8, 16), np.float32)
    r = MeshRenderer(verts, faces, shell_face_uvs(16, 16), tex)
    o = torch.tensor([[5.0, 5.0, 5.0]] * 64).cuda()
    d = torch.nn.functional.normalize(torch.tensor([[1.0, 1.0, 1.0]] * 64), dim=1).cuda()   # pointing away from the mesh
    out = r.render_rays(o, d)["renders"]["ray_traced"]
    assert float(out["is_hit"].sum()) == 0 and float(out["alpha"].abs().sum()) == 0
    assert torch.equal(out["rgb"], torch.ones(64, 3).cuda())             # white background (mesh_renderer.py:55-60)
    with pytest.raises(AssertionError):
        MeshRenderer(verts, faces, shell_face_uvs(16, 16), np.zeros((8, 8, 10), np.float32))
