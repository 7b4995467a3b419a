"""GPU parity of the K-layer shell intersector against the C oracle: hit triangle indices bit-exact, depth / barycentric / position /
normal bit-exact (same IEEE fp32 operation order, no contraction), layer order and packing offsets bit-exact."""
import numpy as np
import pytest
import torch

from oracle.packing import pack_layer_hits as pack_oracle
from oracle.raytrace import OracleRayTracer
from volsurfs_b200.synthetic import camera_rays, shell_meshes

pytestmark = pytest.mark.gpu


def _cmp_layers(got, want):
    depth = got["depth"].cpu().numpy()
    assert np.array_equal(depth, want["depth"]), f"{(depth != want['depth']).sum()} depth mismatches"
    assert np.array_equal(got["tri"].cpu().numpy(), want["tri"])
    assert np.array_equal(got["u"].cpu().numpy(), want["u"])
    assert np.array_equal(got["v"].cpu().numpy(), want["v"])


@pytest.mark.parametrize("shuffle", [None, 7])
def test_shells_vs_bvh_oracle_64k_rays(shuffle):
    """C2 geometry (5 shells x ~100k triangles), 256x256 camera rays, vs the reference-faithful BVH oracle"""
    from volsurfs_b200.raytracer import ShellTracer

    meshes = shell_meshes(K=5)
    o, d = camera_rays(256, 256, shuffle_seed=shuffle)
    want = OracleRayTracer(meshes).trace_layers(o.numpy(), d.numpy(), mode="bvh")
    tracer = ShellTracer(meshes)
    got = tracer.trace_layers(o.cuda(), d.cuda())
    _cmp_layers(got, want)
    assert not tracer.overflowed()
    assert 0.1 < want["is_hit"].mean() < 0.9


def test_shells_vs_brute_force_4k_rays():
    from volsurfs_b200.raytracer import ShellTracer

    meshes = shell_meshes(K=5)
    o, d = camera_rays(64, 64)
    want = OracleRayTracer(meshes).trace_layers(o.numpy(), d.numpy(), mode="brute")
    got = ShellTracer(meshes).trace_layers(o.cuda(), d.cuda())
    _cmp_layers(got, want)


def test_random_rays_from_inside_and_outside():
    """rays starting between the shells, inside the innermost and outside, random directions; 9 layers"""
    from volsurfs_b200.raytracer import ShellTracer

    meshes = shell_meshes(K=9, n_lat=64, n_lon=64, offset=0.006)
    rng = np.random.default_rng(11)
    n = 20000
    o = (rng.standard_normal((n, 3)) * 0.25).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:100, 0] = 0.0  # axis-aligned components (slab test divides by zero)
    d[100:200, 1] = 0.0
    want = OracleRayTracer(meshes).trace_layers(o, d, mode="brute")
    got = ShellTracer(meshes).trace_layers(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    _cmp_layers(got, want)


def test_reference_compatible_trace_dict():
    """ShellTracer.trace(mesh_id) returns the keys / dtypes / values of raytracelib.RayTracer.trace (raytracer.py:103-113)"""
    from volsurfs_b200.raytracer import ShellTracer

    meshes = shell_meshes(K=3, n_lat=96, n_lon=96)
    o, d = camera_rays(128, 128)
    oracle = OracleRayTracer(meshes)
    tracer = ShellTracer(meshes)
    for k in range(3):
        want = oracle.trace(o.numpy(), d.numpy(), k, mode="bvh")
        got = tracer.trace(o.cuda(), d.cuda(), mesh_id=k)
        assert got["triangles_id"].dtype == torch.int64 and got["triangles_mesh_id"].dtype == torch.int64
        assert bool(got["any_hit"]) == want["any_hit"]
        for key in ("is_hit", "positions", "triangles_mesh_id", "triangles_id", "depth", "normals", "barycentric", "view_dirs"):
            assert np.array_equal(got[key].cpu().numpy(), want[key]), key


def test_trace_and_pack_end_to_end_bit_exact():
    """trace -> pack: ray_start_end_idx, samples_idx, layer order, z, positions and face normals vs oracle trace + oracle packing"""
    from volsurfs_b200.raytracer import ShellTracer

    meshes = shell_meshes(K=5, n_lat=128, n_lon=128)
    o, d = camera_rays(200, 200)
    K = 5
    want_l = OracleRayTracer(meshes).trace_layers(o.numpy(), d.numpy(), mode="bvh")
    unc, layer_of_slot = pack_oracle(o.numpy(), d.numpy(), want_l["is_hit"].T, want_l["depth"].T)
    want = unc.compact_to_valid_samples()
    for exact in (True, False):
        rsp = ShellTracer(meshes).render_samples(o.cuda(), d.cuda(), exact_size=exact)
        S = want.get_total_nr_samples()
        assert int(rsp.total_dev.item()) == S
        assert np.array_equal(rsp.ray_start_end_idx.cpu().numpy(), want.ray_start_end_idx)
        for name in ("samples_idx", "samples_z", "samples_3d", "samples_dirs"):
            assert np.array_equal(getattr(rsp, name)[:S].cpu().numpy(), getattr(want, name)), name
        lay = layer_of_slot[want.samples_idx[:, 0]]
        ray = want.samples_idx[:, 0] // K
        assert np.array_equal(rsp.samples_layer[:S].cpu().numpy(), lay)
        assert np.array_equal(rsp.samples_triangle[:S].cpu().numpy(), want_l["tri"][lay, ray])
        normals = np.stack([want_l["per_mesh"][k]["normals"] for k in range(K)])
        assert np.array_equal(rsp.samples_normals[:S].cpu().numpy(), normals[lay, ray])
        # outer -> inner: z increases along each ray's packed segment
        z = rsp.samples_z[:S, 0].cpu().numpy()
        se = want.ray_start_end_idx
        multi = np.nonzero(se[:, 1] - se[:, 0] > 1)[0][:2000]
        assert all(np.all(np.diff(z[se[r, 0]:se[r, 1]]) > 0) for r in multi)
