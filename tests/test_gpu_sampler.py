"""GPU parity of the ray samplers and occupancy-grid queries through the PyBridge-shaped shim / C ABI:

  * against the numpy restatement (oracle/sampler.py) on small cases ............................ bit-exact, every field
  * against the REFERENCE'S OWN KERNELS (oracle/_ref/libsampler_ref.so = RaySamplerGPU.cuh + OccupancyGridGPU.cuh compiled unmodified
    behind oracle/ref_sampler_harness.cu) at 20k rays x 64^3 voxels, compacted like the reference does ......... bit-exact, every field

Index work (ray_start_end_idx, samples_idx) and fp32 sample positions / depths are all held to bit equality: the product follows the
reference's operation order including the contractions nvcc applies to it."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import sampler as osamp
from oracle.importance import PCG_DEFAULT_INC, PCG_DEFAULT_STATE, Pcg32
from sampler_scene import make_scene

pytestmark = pytest.mark.gpu
REF_PATH = ROOT / "oracle" / "_ref" / "libsampler_ref.so"
FIELDS = ("ray_start_end_idx", "samples_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt", "ray_max_dt")


def _reset_rng():
    from volsurfs_b200.volsurfs import RaySampler

    RaySampler._rng_state, RaySampler._rng_inc = PCG_DEFAULT_STATE, PCG_DEFAULT_INC


def _cuda_scene(sc):
    t = {k: torch.from_numpy(np.ascontiguousarray(sc[k])).cuda() for k in ("o", "d", "t_entry", "t_exit", "occ", "roi", "vals")}
    return t


def _assert_packet_equal(got, want, what):
    for k in FIELDS:
        g = getattr(got, k).cpu().numpy() if not isinstance(got, dict) else got[k]
        w = want[k]
        assert g.shape == w.shape, (what, k, g.shape, w.shape)
        assert np.array_equal(g, w), (what, k, int((g != w).sum()), "of", g.size)


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("use_grid", [False, True])
def test_fg_samplers_match_restatement(use_grid, jitter):
    from volsurfs_b200.volsurfs import RaySampler

    sc = make_scene(300, 16, seed=11)
    t = _cuda_scene(sc)
    _reset_rng()
    min_dist, min_nr, max_nr = 0.03, 2, 48
    if use_grid:
        got = RaySampler.compute_samples_fg_in_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"], min_dist, min_nr, max_nr, jitter,
                                                                    sc["n"], sc["extent"], t["occ"], t["roi"], 1)
        grid = osamp.Grid(sc["n"], sc["extent"], sc["occ"], sc["roi"])
    else:
        got = RaySampler.compute_samples_fg(t["o"], t["d"], t["t_entry"], t["t_exit"], min_dist, min_nr, max_nr, jitter, 1)
        grid = None
    want = osamp.compact(osamp.samples_fg(sc["o"], sc["d"], sc["t_entry"], sc["t_exit"], min_dist, min_nr, max_nr, jitter=jitter, rng=Pcg32(),
                                          grid=grid))
    assert got.is_compacted and got.get_total_nr_samples() == want["samples_z"].shape[0] > 100
    _assert_packet_equal(got, want, f"grid={use_grid} jitter={jitter}")
    assert np.array_equal(got.ray_o.cpu().numpy(), sc["o"]) and np.array_equal(got.ray_exit.cpu().numpy(), sc["t_exit"])
    assert got.samples_values.shape == (got.get_total_nr_samples(), 1) and float(got.samples_values.max()) == -1.0
    # the static generator moves on by 2^32 after a jittered call (RaySampler.cu:228-231)
    g = Pcg32()
    if jitter:
        g.advance(1 << 32)
    assert RaySampler._rng_state == g.state


@pytest.mark.parametrize("jitter", [False, True])
def test_bg_sampler_matches_restatement(jitter):
    from volsurfs_b200.volsurfs import RaySampler

    sc = make_scene(200, 16, seed=12)
    t = _cuda_scene(sc)
    _reset_rng()
    got = RaySampler.compute_samples_bg(t["o"], t["d"], t["t_exit"], 40.0, 24, jitter)
    want = osamp.samples_bg(sc["o"], sc["d"], sc["t_exit"], 40.0, 24, jitter=jitter, rng=Pcg32())
    for k in ("ray_start_end_idx", "samples_3d", "samples_dirs", "samples_z", "ray_max_dt"):
        assert np.array_equal(getattr(got, k).cpu().numpy(), want[k]), k
    assert float(got.ray_exit.min()) == 40.0 and got.get_values_dim() == 0


def test_occupancy_grid_queries_match_restatement():
    from volsurfs_b200.volsurfs import OccupancyGrid

    sc = make_scene(300, 16, seed=13)
    t = _cuda_scene(sc)
    og = OccupancyGrid(sc["n"], sc["extent"])
    assert og.get_nr_voxels() == 16 ** 3 and og.get_nr_occupied_voxels() == 16 ** 3
    og.set_grid_occupancy(t["occ"])
    og.set_grid_roi(t["roi"])
    og.set_grid_values(t["vals"])
    assert og.get_nr_occupied_voxels_in_roi() == int((sc["occ"] & sc["roi"]).sum())
    grid = osamp.Grid(sc["n"], sc["extent"], sc["occ"], sc["roi"], sc["vals"])
    near, far = og.get_rays_t_near_t_far(t["o"], t["d"], t["t_entry"], t["t_exit"])
    wn, wf = osamp.rays_t_near_t_far(sc["o"], sc["d"], sc["t_entry"], sc["t_exit"], grid)
    assert np.array_equal(near.cpu().numpy(), wn) and np.array_equal(far.cpu().numpy(), wf)
    pts = (np.random.RandomState(1).rand(2000, 3).astype(np.float32) - 0.5) * 1.6
    occ, val = og.check_occupancy(torch.from_numpy(pts).cuda())
    wo, wv = osamp.check_occupancy(pts, grid)
    assert occ.dtype == torch.bool and np.array_equal(occ.cpu().numpy(), wo) and np.array_equal(val.cpu().numpy(), wv)
    with pytest.raises(RuntimeError):
        OccupancyGrid(12, [1, 1, 1])


# ---- the reference's own kernels ---------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref():
    if not REF_PATH.exists():
        pytest.skip("oracle/_ref/libsampler_ref.so not built (python -m oracle.build where /root/reference is mounted)")
    lib = ctypes.CDLL(str(REF_PATH))
    assert lib.ref_sampler_abi_version() == 2
    return lib


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def _ref_fg(ref, sc, t, min_dist, min_nr, max_nr, jitter, use_grid):
    """the reference kernel into an uncompacted packet (constructor fills of src/RaySamplesPacked.cu:13-48), then compacted"""
    from volsurfs_b200.volsurfs import RaySamplesPacked

    n = sc["o"].shape[0]
    unc = RaySamplesPacked(n, n * max_nr, 0, 1)
    unc.is_compacted = False
    unc.ray_o, unc.ray_d, unc.ray_enter, unc.ray_exit = t["o"].clone(), t["d"].clone(), t["t_entry"].clone(), t["t_exit"].clone()
    common = [P(t["o"]), P(t["d"]), P(t["t_entry"]), P(t["t_exit"]), ctypes.c_float(min_dist), min_nr, max_nr, ctypes.c_uint64(PCG_DEFAULT_STATE),
              ctypes.c_uint64(PCG_DEFAULT_INC), int(jitter)]
    outs = [P(unc.ray_max_dt), P(unc.samples_idx), P(unc.samples_3d), P(unc.samples_dirs), P(unc.samples_z), P(unc.samples_dt),
            P(unc.ray_start_end_idx), n]
    if use_grid:
        ext = (ctypes.c_float * 3)(*[float(v) for v in sc["extent"]])
        code = ref.ref_samples_fg_occupied(*common, sc["n"], ext, P(t["occ"]), P(t["roi"]), *outs)
    else:
        code = ref.ref_samples_fg(*common, *outs)
    assert code == 0, f"reference harness returned CUDA error {code}"
    from test_gpu_packing import ref_compact

    out = ref_compact(ref, unc)   # the reference's own compaction kernel (RaySamplesPackedGPU.cuh:172-257), not the product's
    out.ray_o, out.ray_d, out.ray_enter, out.ray_exit, out.ray_max_dt = unc.ray_o, unc.ray_d, unc.ray_enter, unc.ray_exit, unc.ray_max_dt
    out.is_compacted = True
    return out


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("use_grid", [False, True])
def test_fg_samplers_match_reference_kernels(ref, use_grid, jitter):
    from volsurfs_b200.volsurfs import RaySampler

    sc = make_scene(20000, 64, seed=21)
    t = _cuda_scene(sc)
    min_dist, min_nr, max_nr = 0.004, 1, 256
    want = _ref_fg(ref, sc, t, min_dist, min_nr, max_nr, jitter, use_grid)
    _reset_rng()
    if use_grid:
        got = RaySampler.compute_samples_fg_in_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"], min_dist, min_nr, max_nr, jitter,
                                                                    sc["n"], sc["extent"], t["occ"], t["roi"], 1)
    else:
        got = RaySampler.compute_samples_fg(t["o"], t["d"], t["t_entry"], t["t_exit"], min_dist, min_nr, max_nr, jitter, 1)
    S = want.get_total_nr_samples()
    print(f"grid={use_grid} jitter={jitter}: {S} samples over {int((want.get_nr_samples_per_ray() > 0).sum())} rays")
    assert S > 100000 and got.get_total_nr_samples() == S
    for k in FIELDS:
        assert torch.equal(getattr(got, k), getattr(want, k)), (k, int((getattr(got, k) != getattr(want, k)).sum()))


def test_bg_and_grid_queries_match_reference_kernels(ref):
    from volsurfs_b200.volsurfs import OccupancyGrid, RaySampler

    sc = make_scene(20000, 64, seed=22)
    t = _cuda_scene(sc)
    n, nr = 20000, 32
    ext = (ctypes.c_float * 3)(*[float(v) for v in sc["extent"]])
    for jitter in (False, True):
        z = torch.empty(n * nr, 1, device="cuda")
        p3, dr = torch.empty(n * nr, 3, device="cuda"), torch.empty(n * nr, 3, device="cuda")
        dt, mdt, se = torch.empty(n * nr, 1, device="cuda"), torch.empty(n, 1, device="cuda"), torch.empty(n, 2, dtype=torch.int32, device="cuda")
        assert ref.ref_samples_bg(P(t["o"]), P(t["d"]), P(t["t_exit"]), ctypes.c_float(30.0), nr, ctypes.c_uint64(PCG_DEFAULT_STATE),
                                  ctypes.c_uint64(PCG_DEFAULT_INC), int(jitter), P(mdt), P(p3), P(dr), P(z), P(dt), P(se), n) == 0
        _reset_rng()
        got = RaySampler.compute_samples_bg(t["o"], t["d"], t["t_exit"], 30.0, nr, jitter)
        assert torch.equal(got.samples_z, z) and torch.equal(got.samples_3d, p3) and torch.equal(got.samples_dirs, dr)
        assert torch.equal(got.ray_max_dt, mdt) and torch.equal(got.ray_start_end_idx, se)
    og = OccupancyGrid(sc["n"], sc["extent"])
    og.set_grid_occupancy(t["occ"])
    og.set_grid_roi(t["roi"])
    og.set_grid_values(t["vals"])
    near, far = og.get_rays_t_near_t_far(t["o"], t["d"], t["t_entry"], t["t_exit"])
    rn, rf = torch.empty_like(near), torch.empty_like(far)
    assert ref.ref_rays_t_near_t_far(P(t["o"]), P(t["d"]), P(t["t_entry"]), P(t["t_exit"]), sc["n"], ext, P(t["occ"]), P(t["roi"]), P(rn), P(rf), n) == 0
    assert torch.equal(near, rn) and torch.equal(far, rf)
    pts = ((torch.rand(50000, 3, device="cuda") - 0.5) * 1.5).contiguous()
    occ, val = og.check_occupancy(pts)
    ro, rv = torch.ones(50000, 1, dtype=torch.bool, device="cuda"), torch.ones(50000, 1, device="cuda")
    assert ref.ref_check_occupancy(P(pts), sc["n"], ext, P(t["vals"]), P(t["occ"]), P(t["roi"]), P(ro), P(rv), 50000) == 0
    assert torch.equal(occ, ro) and torch.equal(val, rv)


def _bg_packet(n, nr, t_far, seed):
    from volsurfs_b200.volsurfs import RaySampler

    sc = make_scene(n, 16, seed=seed)
    t = _cuda_scene(sc)
    _reset_rng()
    return RaySampler.compute_samples_bg(t["o"], t["d"], t["t_exit"], t_far, nr, True)


def test_contract_samples_match_restatement_and_reference_kernels(ref):
    """RaySampler.contract_samples / uncontract_samples (src/RaySampler.cu:336-427) on a background packet whose samples straddle the
    |2x| = 1 sphere: bit-exact against the restatement AND the reference's kernels, dt as the reference's closing update_dt(true)"""
    from volsurfs_b200.volsurfs import RaySampler

    n, nr = 3000, 24
    rsp = _bg_packet(n, nr, 40.0, 31)
    rsp.update_dt(True)
    norm2 = (2 * rsp.samples_3d).norm(dim=1)
    assert int((norm2 > 1).sum()) > n and int((norm2 <= 1).sum()) > 0  # both branches
    got_c = RaySampler.contract_samples(rsp)
    got_u = RaySampler.uncontract_samples(got_c)
    o, se = rsp.ray_o.cpu().numpy(), rsp.ray_start_end_idx.cpu().numpy()
    src = rsp
    for unc, got in ((0, got_c), (1, got_u)):
        w3, wz = osamp.contract_samples(o, se, src.samples_3d.cpu().numpy(), src.samples_z.cpu().numpy(), uncontract=bool(unc))
        assert np.array_equal(got.samples_3d.cpu().numpy(), w3) and np.array_equal(got.samples_z.cpu().numpy(), wz), unc
        r3, rz = torch.full_like(src.samples_3d, -7.0), torch.full_like(src.samples_z, -7.0)
        assert ref.ref_contract_samples(P(src.ray_o), P(src.ray_start_end_idx), P(src.samples_3d), P(src.samples_z), P(r3), P(rz), n, n * nr,
                                        unc) == 0
        assert torch.equal(got.samples_3d, r3) and torch.equal(got.samples_z, rz), unc
        # the rest of the packet is a copy; dt re-derived from the new depths as for a background packet
        for k in ("ray_start_end_idx", "samples_dirs", "ray_o", "ray_d", "ray_enter", "ray_exit", "ray_max_dt"):
            assert torch.equal(getattr(got, k), getattr(src, k)), k
        want_dt = src.copy()
        want_dt.samples_z = got.samples_z.clone()
        want_dt.update_dt(True)
        assert got.has_dt and torch.equal(got.samples_dt, want_dt.samples_dt)
        assert got.samples_3d.data_ptr() != src.samples_3d.data_ptr()
        src = got
    # contraction lands inside the unit ball of the scaled coordinates (radius 1 in x, i.e. |2x'| < 2) and the round trip returns
    assert float((2 * got_c.samples_3d).norm(dim=1).max()) < 2.0
    far = (2 * got_c.samples_3d).norm(dim=1) < 1.9  # 1 / (2 - |2x|) amplifies rounding next to the rim
    assert torch.allclose(got_u.samples_3d[far], rsp.samples_3d[far], rtol=1e-4, atol=1e-6)


def test_contract_samples_ragged_and_errors():
    from volsurfs_b200.volsurfs import RaySampler, RaySamplesPacked

    # ragged packet with empty rays: rays of 0, 1, 33 and 70 samples (more than one warp pass)
    counts = np.array([0, 1, 33, 0, 70, 5], np.int32)
    ends = np.cumsum(counts).astype(np.int32)
    se = np.stack([ends - counts, ends], 1).astype(np.int32)
    tot = int(ends[-1])
    rs = np.random.RandomState(5)
    rsp = RaySamplesPacked(len(counts), tot, 0, 1)
    rsp.ray_start_end_idx = torch.from_numpy(se).cuda()
    rsp.ray_o = torch.from_numpy((rs.randn(len(counts), 3) * 0.2).astype(np.float32)).cuda()
    rsp.samples_3d = torch.from_numpy((rs.randn(tot, 3) * 1.5).astype(np.float32)).cuda()
    rsp.samples_z = torch.from_numpy(np.sort(rs.rand(tot, 1).astype(np.float32) * 9, axis=0)).cuda()
    rsp.ray_max_dt = torch.full((len(counts), 1), 0.5, device="cuda")
    rsp.is_compacted = True
    got = RaySampler.contract_samples(rsp)
    w3, wz = osamp.contract_samples(rsp.ray_o.cpu().numpy(), se, rsp.samples_3d.cpu().numpy(), rsp.samples_z.cpu().numpy())
    assert np.array_equal(got.samples_3d.cpu().numpy(), w3) and np.array_equal(got.samples_z.cpu().numpy(), wz)
    back = RaySampler.uncontract_samples(got)
    w3, wz = osamp.contract_samples(rsp.ray_o.cpu().numpy(), se, w3, wz, uncontract=True)
    assert np.array_equal(back.samples_3d.cpu().numpy(), w3) and np.array_equal(back.samples_z.cpu().numpy(), wz)
    # error behaviour of the reference's CHECKs (src/RaySampler.cu:342-343, 389-390)
    rsp.is_compacted = False
    with pytest.raises(RuntimeError, match="compacted"):
        RaySampler.contract_samples(rsp)
    with pytest.raises(RuntimeError, match="compacted"):
        RaySampler.uncontract_samples(rsp)
    with pytest.raises(RuntimeError, match="empty"):
        RaySampler.contract_samples(RaySamplesPacked(0, 0, 0, 1))


def test_grid_maintenance_matches_restatement_and_reference_kernels(ref):
    """voxel sample points, update_grid_values, update_grid_occupancy_with_density_values (src/OccupancyGrid.cu:206-347,446-503):
    bit-exact against the reference's kernels at 64^3 voxels and against the restatement on a subset"""
    from volsurfs_b200.volsurfs import OccupancyGrid

    n, extent = 64, [1.0, 1.2, 0.9]
    ext = (ctypes.c_float * 3)(*extent)
    og = OccupancyGrid(n, extent)
    OccupancyGrid._rng_state, OccupancyGrid._rng_inc = PCG_DEFAULT_STATE, PCG_DEFAULT_INC
    V = n ** 3
    ll, idx = og.get_grid_lower_left_voxels_vertices()
    assert idx.dtype == torch.int32 and torch.equal(idx, torch.arange(V, dtype=torch.int32, device="cuda")) and ll.shape == (V, 3)
    r = torch.empty_like(ll)
    assert ref.ref_grid_points(P(idx), n, ext, 0, ctypes.c_uint64(0), ctypes.c_uint64(1), 0, P(r), V) == 0
    assert torch.equal(ll, r)
    for jitter in (False, True, True):  # twice jittered: the generator advances between calls like the reference's static m_rng
        state = OccupancyGrid._rng_state
        pts, idx2 = og.get_grid_samples(jitter)
        assert ref.ref_grid_points(P(idx2), n, ext, 1, ctypes.c_uint64(state), ctypes.c_uint64(PCG_DEFAULT_INC), int(jitter), P(r), V) == 0
        assert torch.equal(pts, r), jitter
        assert (OccupancyGrid._rng_state != state) == jitter
        sub = slice(0, 3000)
        rng = Pcg32()
        rng.state = state
        want = osamp.grid_points(idx2[sub].cpu().numpy(), n, extent, centre=True, jitter=jitter, rng=rng)
        assert np.array_equal(pts[sub].cpu().numpy(), want), jitter
        occ, _ = og.check_occupancy(pts)  # every sample lies in its own voxel: all inside the (full) grid
        assert bool(occ.all())
    state = OccupancyGrid._rng_state
    pts, ridx = og.get_random_grid_samples(5000, True)
    assert ridx.shape == (5000,) and ridx.dtype == torch.int32 and int(ridx.min()) >= 0 and int(ridx.max()) < V
    r5 = torch.empty_like(pts)
    assert ref.ref_grid_points(P(ridx), n, ext, 1, ctypes.c_uint64(state), ctypes.c_uint64(PCG_DEFAULT_INC), 1, P(r5), 5000) == 0
    assert torch.equal(pts, r5)
    og.init_sphere_roi(0.5, 0.05)
    roi = og.get_grid_roi()
    assert roi.dtype == torch.bool and 0 < og.get_nr_voxels_in_roi() < V
    corners_far = (ll.norm(dim=1) >= 0.45)
    assert not bool((roi & corners_far).any())  # a voxel whose lower-left corner is outside the sphere is not in the roi
    pts, ridx = og.get_random_grid_samples_in_roi(4000, False)
    assert bool(roi[ridx.long()].all()) and pts.shape == (4000, 3)

    # update_grid_values on unique indices (duplicates race in the reference), then the two occupancy rules
    g = torch.Generator(device="cuda").manual_seed(4)
    vals0 = torch.rand(V, device="cuda", generator=g)
    og.set_grid_values(vals0.clone())
    pick = torch.randperm(V, device="cuda", generator=g)[:100000].to(torch.int32)
    new = torch.rand(100000, 1, device="cuda", generator=g)
    og.update_grid_values(pick, new, 0.95)
    rv = vals0.clone()
    assert ref.ref_update_grid_values(P(pick), P(new), ctypes.c_float(0.95), n, P(rv), 100000) == 0
    assert torch.equal(og.get_grid_values(), rv)
    assert np.array_equal(osamp.update_grid_values(pick.cpu().numpy(), new.cpu().numpy(), 0.95, vals0.cpu().numpy()), rv.cpu().numpy())
    assert og.get_grid_max_value_in_roi() <= og.get_grid_max_value() and og.get_grid_min_value_in_roi() >= og.get_grid_min_value()
    sparse = torch.where(torch.rand(V, device="cuda", generator=g) < 0.02, vals0, torch.zeros_like(vals0))  # isolated dense voxels
    og.set_grid_values(sparse)
    for unit_extent in (True, False):
        e = [1.0, 1.0, 1.0] if unit_extent else extent
        og2 = OccupancyGrid(n, e)
        og2.set_grid_values(sparse)
        ec = (ctypes.c_float * 3)(*e)
        for neigh in (False, True):
            og2.set_grid_occupancy(torch.rand(V, device="cuda", generator=g) < 0.5)
            ro = og2.get_grid_occupancy().clone()
            og2.update_grid_occupancy_with_density_values(pick, 0.3, neigh)
            assert ref.ref_update_grid_occupancy_density(P(pick), n, ec, ctypes.c_float(0.3), int(neigh), P(sparse), P(ro), 100000) == 0
            assert torch.equal(og2.get_grid_occupancy(), ro), (unit_extent, neigh)
            sub = pick[:1500]
            before = og2.get_grid_occupancy().clone()
            want = osamp.update_grid_occupancy_density(sub.cpu().numpy(), n, e, 0.3, neigh, sparse.cpu().numpy(), before.cpu().numpy())
            og2.update_grid_occupancy_with_density_values(sub, 0.3, neigh)
            assert np.array_equal(og2.get_grid_occupancy().cpu().numpy(), want), (unit_extent, neigh)
    with pytest.raises(RuntimeError):
        og.update_grid_values(pick, new.reshape(-1), 0.95)
    with pytest.raises(RuntimeError):
        og.update_grid_values(pick, new, 1.5)
    with pytest.raises(RuntimeError):
        og.update_grid_occupancy_with_density_values(pick.reshape(-1, 1), 0.3, False)


def test_sdf_occupancy_rule_matches_reference_kernel_and_restatement(ref):
    """OccupancyGrid.update_grid_occupancy_with_sdf_values (src/OccupancyGrid.cu:505-533) as surf.py:297 calls it: grid values = SDF at
    the voxel centres, one beta per index; bit-exact against the reference's kernel, restatement equal away from the threshold"""
    from volsurfs_b200.volsurfs import OccupancyGrid

    n, extent = 64, [1.0, 1.2, 0.9]
    V = n ** 3
    ext = (ctypes.c_float * 3)(*extent)
    og = OccupancyGrid(n, extent)
    pts, idx = og.get_grid_samples(False)
    sdf = (pts.norm(dim=1) - 0.3).contiguous()  # a sphere of radius 0.3
    og.update_grid_values(idx, sdf.reshape(-1, 1), 0.0)  # decay 0: max(new, 0 * old) keeps negative sdf at 0 ...
    og.set_grid_values(sdf.clone())                     # ... so the methods write the SDF itself, like surf.py does after its own update
    g = torch.Generator(device="cuda").manual_seed(9)
    for beta_scale, thresh in ((50.0, 1e-4), (400.0, 1e-2)):
        beta = (beta_scale * (0.5 + torch.rand(V, 1, device="cuda", generator=g))).contiguous()
        og.set_grid_occupancy(torch.rand(V, device="cuda", generator=g) < 0.5)
        ro = og.get_grid_occupancy().clone()
        before = ro.clone()
        og.update_grid_occupancy_with_sdf_values(idx, beta, thresh, False)
        assert ref.ref_update_grid_occupancy_sdf(P(idx), n, ext, P(beta), ctypes.c_float(thresh), 0, P(sdf), P(ro), V) == 0
        got = og.get_grid_occupancy()
        assert torch.equal(got, ro), (beta_scale, int((got != ro).sum()))
        assert 0 < int(got.sum()) < V  # a shell around the sphere
        want, w = osamp.update_grid_occupancy_sdf(idx.cpu().numpy(), n, extent, beta.cpu().numpy(), thresh, sdf.cpu().numpy(),
                                                  before.cpu().numpy(), return_weight=True)
        clear = np.abs(w - np.float32(thresh)) > 1e-4 * thresh
        assert clear.mean() > 0.999 and np.array_equal(got.cpu().numpy()[clear], want[clear])
    # a subset of indices rewrites only those voxels
    og.set_grid_occupancy_full()
    sub = idx[::7].contiguous()
    og.update_grid_occupancy_with_sdf_values(sub, torch.full((sub.shape[0], 1), 200.0, device="cuda"), 1e-3, True)
    occ = og.get_grid_occupancy()
    mask = torch.zeros(V, dtype=torch.bool, device="cuda")
    mask[sub.long()] = True
    assert bool(occ[~mask].all()) and not bool(occ[mask].all())
    with pytest.raises(RuntimeError):
        og.update_grid_occupancy_with_sdf_values(sub, torch.full((3, 1), 200.0, device="cuda"), 1e-3, True)


def test_first_sample_start_matches_reference_kernel(ref):
    """OccupancyGrid.get_first_rays_sample_start_of_grid_occupied_regions (src/OccupancyGrid.cu:536-573, called by
    utils/sphere_tracing.py:42): every field of the one-sample-per-ray packet bit-exact against the reference's kernel"""
    from volsurfs_b200.volsurfs import OccupancyGrid, RaySamplesPacked

    sc = make_scene(20000, 64, seed=41)
    t = _cuda_scene(sc)
    n = 20000
    og = OccupancyGrid(sc["n"], sc["extent"])
    og.set_grid_occupancy(t["occ"])
    og.set_grid_roi(t["roi"])
    got = og.get_first_rays_sample_start_of_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"])
    want = RaySamplesPacked(n, n, 0, 1)  # the reference's constructor fills (src/RaySamplesPacked.cu:13-48)
    ext = (ctypes.c_float * 3)(*[float(v) for v in sc["extent"]])
    assert ref.ref_first_sample_start(P(t["o"]), P(t["d"]), P(t["t_entry"]), P(t["t_exit"]), sc["n"], ext, P(t["occ"]), P(t["roi"]),
                                      P(want.samples_3d), P(want.samples_dirs), P(want.samples_z), P(want.samples_dt),
                                      P(want.ray_start_end_idx), n) == 0
    for k in ("ray_start_end_idx", "samples_3d", "samples_dirs", "samples_z", "samples_dt"):
        assert torch.equal(getattr(got, k), getattr(want, k)), k
    se = got.ray_start_end_idx
    hit = se[:, 1] > se[:, 0]
    assert 0 < int(hit.sum()) < n
    rows = torch.arange(n, dtype=torch.int32, device="cuda")
    assert torch.equal(se[hit, 0], rows[hit]) and torch.equal(se[hit, 1], rows[hit] + 1) and not bool(se[~hit].any())
    occ, _ = og.check_occupancy(got.samples_3d[hit])
    assert bool(occ.all())  # the stored position lies in an occupied voxel of the roi
    near, _ = og.get_rays_t_near_t_far(t["o"], t["d"], t["t_entry"], t["t_exit"])
    assert bool((got.samples_z[hit] > near[hit]).all())  # depth = t after stepping out of that voxel
    assert bool((got.samples_3d[~hit] == -1).all())  # rays without a hit keep the constructor's fill


def test_advance_to_next_occupied_voxel_matches_reference_kernel(ref):
    """OccupancyGrid.advance_ray_sample_to_next_occupied_voxel (src/OccupancyGrid.cu:575-607): bit-exact against the reference's kernel
    for points that leave the grid through an upper face (the reference's kernel does not return for the others: its index clamps
    coordinates below the grid to voxel 0); the input positions are updated in place, as the reference does"""
    from volsurfs_b200.volsurfs import OccupancyGrid

    sc = make_scene(64, 64, seed=43)
    t = _cuda_scene(sc)
    n = 20000
    og = OccupancyGrid(sc["n"], sc["extent"])
    og.set_grid_occupancy(t["occ"])
    og.set_grid_roi(t["roi"])
    g = torch.Generator(device="cuda").manual_seed(11)
    extent = torch.tensor([float(v) for v in sc["extent"]], device="cuda")
    start = ((torch.rand(n, 3, device="cuda", generator=g) - 0.5) * 0.9 * extent).contiguous()  # inside the grid
    dirs = torch.nn.functional.normalize(torch.rand(n, 3, device="cuda", generator=g) + 0.05, dim=1).contiguous()  # all components > 0
    start[::50] = 5.0  # some points outside the grid: returned unchanged, not within bounds
    ext = (ctypes.c_float * 3)(*[float(v) for v in sc["extent"]])
    r3, rw = torch.full_like(start, -7.0), torch.ones(n, 1, dtype=torch.bool, device="cuda")
    assert ref.ref_advance_to_next_occupied(P(dirs), P(start), sc["n"], ext, P(t["occ"]), P(t["roi"]), P(r3), P(rw), n) == 0
    src = start.clone()
    got, within = og.advance_ray_sample_to_next_occupied_voxel(dirs, src)
    assert got.data_ptr() == src.data_ptr() and within.dtype == torch.bool and within.shape == (n, 1)
    assert torch.equal(got, r3) and torch.equal(within, rw)
    w = within[:, 0]
    assert 0 < int(w.sum()) < n and not bool(w[::50].any()) and torch.equal(got[::50], start[::50])
    occ, _ = og.check_occupancy(got[w])
    assert bool(occ.all())  # points that stayed inside stopped in an occupied voxel of the roi
    # points leaving through a lower face end the march too (deviation: the reference never returns for them)
    src = start[1:2000:2].clone()
    got, within = og.advance_ray_sample_to_next_occupied_voxel(-dirs[1:2000:2].contiguous(), src)
    w = within[:, 0]
    assert 0 < int(w.sum()) < w.numel()
    occ, _ = og.check_occupancy(got[w])
    assert bool(occ.all()) and bool((got[~w].abs() <= 0.5 * extent + 1e-4).all())  # the others: last position probed inside the grid
    with pytest.raises(RuntimeError):
        og.advance_ray_sample_to_next_occupied_voxel(dirs[:5], src)


def test_sampler_feeds_packed_compositing():
    """the sampler's packet goes straight into update_dt and the packed operators (the NeRF path of volsurfs_py/methods/nerf.py:280-334)"""
    from volsurfs_b200.volsurfs import RaySampler, VolumeRendering

    sc = make_scene(5000, 32, seed=23)
    t = _cuda_scene(sc)
    _reset_rng()
    rsp = RaySampler.compute_samples_fg_in_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"], 0.01, 1, 128, True, sc["n"], sc["extent"],
                                                                 t["occ"], t["roi"], 1)
    rsp.update_dt(False)
    S = rsp.get_total_nr_samples()
    assert S > 0 and float(rsp.samples_dt.min()) >= 0.0
    sigma = torch.rand(S, 1, device="cuda") * 20
    alpha = 1.0 - torch.exp(-sigma * rsp.samples_dt)
    T, bg = VolumeRendering.cumprod_one_minus_alpha_to_transmittance(rsp, 1.0 - alpha + 1e-6)
    w = alpha * T
    acc = VolumeRendering.integrate_with_weights_1d(rsp, torch.ones_like(w), w)
    assert float(acc.max()) <= 1.0 + 1e-4 and torch.isfinite(acc).all() and bg.shape == (5000, 1)


def test_argument_errors():
    from volsurfs_b200.volsurfs import RaySampler

    o = torch.zeros(4, 3, device="cuda")
    with pytest.raises(RuntimeError):
        RaySampler.compute_samples_fg(o, o[:3], torch.zeros(4, 1, device="cuda"), torch.ones(4, 1, device="cuda"), 0.1, 1, 8, False, 1)
    with pytest.raises(RuntimeError):
        RaySampler.compute_samples_fg_in_grid_occupied_regions(o, o, torch.zeros(4, 1, device="cuda"), torch.ones(4, 1, device="cuda"), 0.1, 1, 8, False,
                                                              8, [1, 1, 1], torch.ones(10, dtype=torch.bool, device="cuda"),
                                                              torch.ones(512, dtype=torch.bool, device="cuda"), 1)


def test_init_with_one_sample_per_ray():
    from volsurfs_b200.volsurfs import RaySampler

    p, d = torch.rand(100, 3, device="cuda"), torch.rand(100, 3, device="cuda")
    rsp = RaySampler.init_with_one_sample_per_ray(p, d)
    assert rsp.get_total_nr_samples() == 100 and torch.equal(rsp.samples_3d, p) and torch.equal(rsp.samples_dirs, d)
    assert float(rsp.samples_z.abs().max()) == 0.0 and float(rsp.samples_dt.abs().max()) == 0.0
    assert rsp.ray_start_end_idx[7].tolist() == [7, 8]


def test_full_size_properties():
    """BASELINE config[2] size (640k rays, 128^3 voxels, <= 1024 samples per ray): size-independent properties of the sampler's packet"""
    from volsurfs_b200.volsurfs import OccupancyGrid, RaySampler

    sc = make_scene(640000, 128, seed=41)
    t = _cuda_scene(sc)
    _reset_rng()
    rsp = RaySampler.compute_samples_fg_in_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"], 0.0015, 1, 1024, True, sc["n"],
                                                                 sc["extent"], t["occ"], t["roi"], 1)
    S = rsp.get_total_nr_samples()
    se = rsp.ray_start_end_idx
    cnt = rsp.get_nr_samples_per_ray().long()
    assert S > 10_000_000 and int(cnt.max()) <= 1024 and int(cnt.sum()) == S
    has = cnt > 0
    # segments are the exclusive prefix sum of the counts, empty rays carry (-1,-1)
    start = torch.cumsum(cnt, 0) - cnt
    assert torch.equal(se[has, 0].long(), start[has]) and torch.equal(se[has, 1].long(), (start + cnt)[has]) and bool((se[~has] == -1).all())
    # depths strictly increase inside a ray and stay inside [t_entry, t_exit]
    ray_of = torch.repeat_interleave(torch.arange(cnt.numel(), device="cuda"), cnt)
    z = rsp.samples_z.view(-1)
    same_ray = ray_of[1:] == ray_of[:-1]
    assert bool((z[1:][same_ray] > z[:-1][same_ray]).all())
    assert bool((z >= t["t_entry"].view(-1)[ray_of]).all()) and bool((z <= t["t_exit"].view(-1)[ray_of]).all())
    # every sample sits in an occupied voxel of the region of interest; positions are o + z d
    og = OccupancyGrid(sc["n"], sc["extent"])
    og.set_grid_occupancy(t["occ"])
    og.set_grid_roi(t["roi"])
    occ, _ = og.check_occupancy(rsp.samples_3d)
    assert bool(occ.all())
    want = torch.addcmul(t["o"][ray_of], z.unsqueeze(1), t["d"][ray_of])
    assert float((rsp.samples_3d - want).abs().max()) < 1e-6
    # samples_idx = slot in the reference's uncompacted packet
    pos_in_ray = torch.arange(S, device="cuda") - start[ray_of]
    assert torch.equal(rsp.samples_idx.view(-1).long(), ray_of * 1024 + pos_in_ray)
