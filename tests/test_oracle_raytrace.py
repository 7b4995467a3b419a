"""CPU tests of the ray-tracing oracle (C restatement of raytracelib): reference-faithful BVH traversal vs brute force in
index order, closed-form intersections, the reference's edge rules."""
import numpy as np
import pytest

from oracle.raytrace import OracleRayTracer
from volsurfs_b200.synthetic import camera_rays, shell_meshes


def test_bvh_equals_brute_force_on_shells():
    meshes = shell_meshes(K=3, n_lat=48, n_lon=48)
    o, d = camera_rays(96, 96)
    rt = OracleRayTracer(meshes)
    a = rt.trace_layers(o.numpy(), d.numpy(), mode="bvh")
    b = rt.trace_layers(o.numpy(), d.numpy(), mode="brute")
    assert np.array_equal(a["depth"], b["depth"])
    assert np.array_equal(a["tri"], b["tri"])
    assert np.array_equal(a["u"], b["u"]) and np.array_equal(a["v"], b["v"])
    assert not any(r["stack_overflow"] for r in a["per_mesh"])
    both = a["is_hit"].all(axis=0)
    assert both.sum() > 500
    # nested shells: outer layers are hit first
    assert np.all(a["depth"][2][both] < a["depth"][1][both]) and np.all(a["depth"][1][both] < a["depth"][0][both])
    for r in a["per_mesh"]:
        h = r["is_hit"]
        assert np.allclose(np.linalg.norm(r["normals"][h], axis=1), 1.0, atol=1e-5)
        assert np.allclose(r["barycentric"][h].sum(1), 1.0, atol=1e-5)
        assert np.all(r["triangles_id"][~h] == -1) and np.all(r["depth"][~h] == np.float32(1e6))
        assert np.all(r["normals"][~h] == 0)


def test_single_triangle_closed_form_and_edge_rules():
    # a fan of 9 copies of one triangle in the plane z=0 (BVH needs > 8 triangles, raytracer.py:17)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    verts = np.concatenate([v + np.array([3 * i, 0, 0], np.float32) for i in range(9)])
    faces = np.arange(27, dtype=np.int32).reshape(9, 3)
    rt = OracleRayTracer([(verts, faces)])
    o = np.array([[0.25, 0.25, 2], [0.25, 0.25, -2], [0.9, 0.9, 2], [0.0, 0.0, 2], [0.25, 0.25, 0.0], [0.25, 0.25, 2]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, -1], [1, 0, 0]], np.float32)
    for mode in ("bvh", "brute"):
        r = rt.trace(o, d, 0, mode)
        assert np.allclose(r["depth"][:2], 2.0)                    # both faces are hit (no back-face culling)
        assert r["triangles_id"][0] == 0 and r["triangles_id"][1] == 0
        assert np.allclose(r["barycentric"][0], [0.5, 0.25, 0.25])  # (1-u-v, u, v), bvh.cu:459
        assert r["depth"][2] == np.float32(1e6) and r["triangles_id"][2] == -1   # u+v > 1
        if mode == "brute":
            assert r["depth"][3] == 2.0                             # vertex hit: u=v=0 is inside (inclusive bounds)
        else:
            # reference artefact kept by the faithful traversal: the ray runs inside the box plane x=0 with d.x == 0, the slab
            # test computes 0/0 = NaN (bounding_box.cuh:152-153), NaN < curr_t is false and the child is never pushed
            assert r["depth"][3] == np.float32(1e6)
        assert r["depth"][4] == np.float32(1e6)                     # origin on the triangle: t == 0 is not > min_t
        assert r["depth"][5] == np.float32(1e6)                     # ray parallel to the plane: D = inf -> rejected
        assert np.array_equal(r["is_hit"], r["depth"] <= 100.0)


# ---- the pin: the reference's own tracer (src/bvh.cu compiled where it lies, oracle/_ref/libraytrace_ref.so) ---------------------------
KEYS = ("depth", "triangles_id", "triangles_mesh_id", "positions", "normals", "barycentric")


def test_host_contract_equals_the_reference_host_path_on_shells():
    """oracle/raytrace_oracle.c (contract "host") against TriangleBvh4::ray_intersect + the raytrace_kernel body run on the CPU:
    same BVH size, every output bit for bit"""
    import pytest

    from oracle import ref_raytrace

    if not ref_raytrace.available():
        pytest.skip("oracle/_ref/libraytrace_ref.so not built (needs /root/reference)")
    meshes = shell_meshes(K=3, n_lat=48, n_lon=48)
    o, d = camera_rays(96, 96)
    rng = np.random.default_rng(5)
    o2 = (rng.standard_normal((4000, 3)) * 0.25).astype(np.float32)      # origins inside / between the shells
    d2 = rng.standard_normal((4000, 3)).astype(np.float32)
    d2[:50, 0] = 0.0                                                      # the slab test divides by zero
    oo = np.concatenate([o.numpy(), o2])
    dd = np.concatenate([d.numpy(), d2])
    oracle = OracleRayTracer(meshes, contract="host")
    ref = ref_raytrace.RefRayTracer(meshes)
    for k in range(3):
        assert oracle.num_nodes(k) == ref.num_nodes(k)
        a = oracle.trace(oo, dd, k)
        b = ref.trace_host(oo, dd, k)
        for key in KEYS:
            assert np.array_equal(a[key], b[key], equal_nan=True), (k, key)
        assert a["is_hit"].sum() > 1000


@pytest.mark.parametrize("name", ["smurf", "plushy"])
def test_reference_meshes_against_reference_run_goldens(name):
    """The reference's own test meshes (raytracelib/meshes/{smurf,plushy}.obj) and probe ray (tests/test_raytracing.py:18-19):
    goldens recorded from the reference's host traversal (tests/golden/make_golden_raytrace.py); the oracle reproduces every bit, the
    brute-force pass agrees wherever no two triangles tie in t"""
    from conftest import GOLDEN

    m = np.load(GOLDEN / f"mesh_{name}.npz")
    g = np.load(GOLDEN / f"raytrace_{name}_host.npz")
    oracle = OracleRayTracer([(m["verts"], m["faces"])], contract="host")
    assert oracle.num_nodes(0) == int(g["nodes"])
    a = oracle.trace(g["rays_o"], g["rays_d"], 0)
    for key in KEYS:
        assert np.array_equal(a[key], g[key], equal_nan=True), key
    assert not a["stack_overflow"]
    assert g["depth"][0] < 100.0 and g["triangles_id"][0] >= 0          # the probe ray (2,0,0) -> (-1,0,0) hits both meshes
    assert (g["depth"] <= 100.0).mean() > 0.3
    b = oracle.trace(g["rays_o"], g["rays_d"], 0, mode="brute")
    same = b["triangles_id"] == g["triangles_id"]
    assert same.mean() > 0.999
    assert np.array_equal(b["depth"][same], g["depth"][same])
    # where brute force (conservative: no boxes) and the reference BVH differ, it is a t-tie between neighbours or a box culled by the
    # reference's non-conservative slab test — never a nearer hit missed by more than rounding
    diff = ~same
    assert np.all(b["depth"][diff] <= g["depth"][diff])


def test_device_contract_differs_only_in_the_last_bits():
    """contract "device" (the reference kernel's FMA contractions) vs "host": same triangles except on near-ties, t within a few ulp"""
    meshes = shell_meshes(K=2, n_lat=48, n_lon=48)
    o, d = camera_rays(96, 96)
    a = OracleRayTracer(meshes, contract="host").trace(o.numpy(), d.numpy(), 1)
    b = OracleRayTracer(meshes, contract="device").trace(o.numpy(), d.numpy(), 1)
    same = a["triangles_id"] == b["triangles_id"]
    assert same.mean() > 0.999
    h = same & a["is_hit"]
    assert np.max(np.abs(a["depth"][h] - b["depth"][h]) / a["depth"][h]) < 1e-5
    assert not np.array_equal(a["depth"], b["depth"])                    # the contraction is visible


def test_packing_oracle_vs_reference_lines_golden():
    """oracle/packing.py:pack_layer_hits (the packed form of the K-layer bookkeeping) against tests/golden/layers_k3.npz, recorded by exec'ing
    volsurfs_py/methods/volsurfs.py:449-516: scattering the packed samples back by (ray, layer) gives the reference's dense buffers"""
    from conftest import GOLDEN
    from oracle.packing import pack_layer_hits

    g = np.load(GOLDEN / "layers_k3.npz")
    K, n_lat, n_lon, _ = (int(x) for x in g["params"])
    meshes = shell_meshes(K=K, n_lat=n_lat, n_lon=n_lon)
    tr = OracleRayTracer(meshes, contract="device").trace_layers(g["rays_o"], g["rays_d"])
    assert np.array_equal(tr["is_hit"].T, g["surfs_hits"])
    unc, layer_of_slot = pack_layer_hits(g["rays_o"], g["rays_d"], tr["is_hit"].T, tr["depth"].T)
    rsp = unc.compact_to_valid_samples()
    ray = rsp.samples_idx[:, 0] // K
    lay = layer_of_slot[rsp.samples_idx[:, 0]]
    N = g["rays_o"].shape[0]
    pts = np.zeros((N, K, 3), np.float32)
    pts[ray, lay] = rsp.samples_3d
    assert np.array_equal(pts, g["surfs_points"])
    nrm = np.zeros((N, K, 3), np.float32)
    normals = np.stack([r["normals"] for r in tr["per_mesh"]])
    nrm[ray, lay] = normals[lay, ray]
    assert np.array_equal(nrm, g["surfs_normals"])
    se = rsp.ray_start_end_idx
    cnt = g["surfs_hits"].sum(1)
    assert np.array_equal(se[:, 1] - se[:, 0], cnt)
    assert np.all(se[cnt == 0] == -1)
