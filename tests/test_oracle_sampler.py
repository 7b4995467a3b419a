"""CPU: known answers and invariants of the sampler / occupancy-grid restatement (oracle/sampler.py).  The reference has no vectors for
these functions; the restatement is pinned against the reference KERNELS on the GPU (tests/test_gpu_sampler.py)."""
import numpy as np

from oracle import sampler as S
from oracle.importance import Pcg32
from sampler_scene import make_scene

F = np.float32


def test_morton_known_answers():
    assert S.morton3d(1, 0, 0) == 1 and S.morton3d(0, 1, 0) == 2 and S.morton3d(0, 0, 1) == 4
    assert S.morton3d(3, 3, 3) == 63 and S.morton3d(2, 0, 0) == 8 and S.morton3d(7, 7, 7) == 511
    # pos_to_lin_idx: grid centred at the origin (occ_grid_helpers.h:55-79)
    assert S.pos_to_lin_idx([F(-0.49), F(-0.49), F(-0.49)], 8, [1, 1, 1]) == 0
    assert S.pos_to_lin_idx([F(0.49), F(0.49), F(0.49)], 8, [1, 1, 1]) == 511
    assert S.pos_to_lin_idx([F(0.01), F(-0.49), F(-0.49)], 8, [1, 1, 1]) == S.morton3d(4, 0, 0)
    assert S.pos_to_lin_idx([F(-5.0), F(0.0), F(0.0)], 8, [1, 1, 1]) == S.morton3d(0, 4, 4)  # negative -> 0 (hardware conversion)


def test_distance_to_next_voxel_is_axis_distance_plus_eps():
    # 8 voxels over extent 1 -> voxel 0.125; at x = 0.01 moving +x the next boundary is at 0.125: 0.115 away along the AXIS
    t = S.distance_to_next_voxel([F(0.01), F(0.0), F(0.0)], [F(1.0), F(0.0), F(0.0)], 8, [1, 1, 1])
    assert abs(float(t) - (0.115 + 1e-6)) < 1e-6
    # not divided by the direction component (reference quirk): a slanted ray gets the same per-axis distances
    t2 = S.distance_to_next_voxel([F(0.01), F(0.0), F(0.0)], [F(0.5), F(0.0), F(0.0)], 8, [1, 1, 1])
    assert t2 == t
    assert S.distance_to_next_voxel([F(0), F(0), F(0)], [F(0), F(0), F(0)], 8, [1, 1, 1]) == F(1e10)


def test_samples_fg_worked_example():
    o, d = np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32)
    out = S.samples_fg(o, d, [[1.0]], [[2.0]], 0.3, 1, 8)
    assert out["ray_start_end_idx"].tolist() == [[0, 3]]  # int(1 / 0.3) = 3 samples, spacing 1/3
    assert np.allclose(out["samples_z"][:3, 0], [1.0, 4 / 3, 5 / 3], atol=1e-6)
    assert np.allclose(out["ray_max_dt"], 1 / 3, atol=1e-7)
    assert out["samples_idx"][:, 0].tolist() == [0, 1, 2, -1, -1, -1, -1, -1]
    # fewer than min_nr samples -> nothing, (-1,-1), ray_max_dt keeps the constructor fill
    out = S.samples_fg(o, d, [[1.0]], [[2.0]], 0.3, 5, 8)
    assert out["ray_start_end_idx"].tolist() == [[-1, -1]] and out["ray_max_dt"][0, 0] == -1 and (out["samples_idx"] == -1).all()
    # entry >= exit -> no samples
    assert S.samples_fg(o, d, [[2.0]], [[1.0]], 0.3, 1, 8)["ray_start_end_idx"].tolist() == [[-1, -1]]
    # jitter shifts the first sample by spacing * U[0,1) drawn from the stream advanced by the ray index
    rng = Pcg32()
    j = S.samples_fg(o, d, [[1.0]], [[2.0]], 0.3, 1, 8, jitter=True, rng=rng)
    g = Pcg32()
    g.advance(0)
    assert np.isclose(j["samples_z"][0, 0], 1.0 + (1 / 3) * float(g.next_float()), atol=1e-6)


def test_samples_fg_in_grid_invariants():
    sc = make_scene(120, 16, seed=3)
    grid = S.Grid(sc["n"], sc["extent"], sc["occ"], sc["roi"], sc["vals"])
    out = S.samples_fg(sc["o"], sc["d"], sc["t_entry"], sc["t_exit"], 0.02, 1, 64, grid=grid)
    se = out["ray_start_end_idx"]
    cnt = se[:, 1] - se[:, 0]
    assert cnt.max() <= 64 and (cnt > 0).sum() > 20 and (cnt == 0).sum() > 5
    for r in np.nonzero(cnt > 0)[0]:
        z = out["samples_z"][se[r, 0]:se[r, 1], 0]
        assert np.all(np.diff(z) > 0) and z[0] >= sc["t_entry"][r, 0] and z[-1] <= sc["t_exit"][r, 0]
        # every sample lies in an occupied voxel of the region of interest
        for p in out["samples_3d"][se[r, 0]:se[r, 1]]:
            assert grid.occupied(S.pos_to_lin_idx(p, grid.n, grid.extent))
    comp = S.compact(out)
    assert comp["samples_z"].shape[0] == cnt.sum()
    assert np.array_equal(comp["ray_start_end_idx"][cnt > 0, 1] - comp["ray_start_end_idx"][cnt > 0, 0], cnt[cnt > 0])
    assert (comp["ray_start_end_idx"][cnt == 0] == -1).all() and (comp["samples_idx"] >= 0).all()
    # an empty grid yields nothing; t_near == t_far == t_entry there
    empty = S.Grid(sc["n"], sc["extent"], np.zeros_like(sc["occ"]), sc["roi"], sc["vals"])
    assert (S.samples_fg(sc["o"], sc["d"], sc["t_entry"], sc["t_exit"], 0.02, 1, 64, grid=empty)["ray_start_end_idx"] == -1).all()
    near, far = S.rays_t_near_t_far(sc["o"][:20], sc["d"][:20], sc["t_entry"][:20], sc["t_exit"][:20], empty)
    assert np.array_equal(near, sc["t_entry"][:20]) and np.array_equal(far, sc["t_entry"][:20])
    near, far = S.rays_t_near_t_far(sc["o"], sc["d"], sc["t_entry"], sc["t_exit"], grid)
    assert np.all(far >= near) and np.all(near >= sc["t_entry"]) and np.all(far <= np.maximum(sc["t_exit"], sc["t_entry"]))


def test_samples_bg_invariants():
    sc = make_scene(10, 16, seed=4)
    out = S.samples_bg(sc["o"], sc["d"], sc["t_exit"], 50.0, 16)
    z = out["samples_z"].reshape(10, 16)
    assert np.allclose(z[:, 0], sc["t_exit"][:, 0], atol=2e-6) and np.all(np.diff(z, axis=1) >= 0) and z.max() <= 50.0
    assert np.allclose(out["ray_max_dt"][:, 0], np.diff(np.concatenate([sc["t_exit"], z], 1), axis=1).max(1), rtol=1e-6)
    j = S.samples_bg(sc["o"], sc["d"], sc["t_exit"], 50.0, 16, jitter=True)
    zj = j["samples_z"].reshape(10, 16)
    assert np.array_equal(zj[:, 0], z[:, 0]) and np.array_equal(zj[:, -1], z[:, -1]) and np.all(zj[:, 1:-1] <= z[:, 1:-1])


def test_check_occupancy():
    sc = make_scene(5, 16, seed=5)
    grid = S.Grid(sc["n"], sc["extent"], sc["occ"], sc["roi"], sc["vals"])
    pts = np.array([[0, 0, 0], [10, 0, 0], [0.4, 0.5, 0.4], [-0.2, 0.1, 0.05]], np.float32)
    occ, val = S.check_occupancy(pts, grid)
    v0 = S.pos_to_lin_idx(pts[0], 16, sc["extent"])
    assert occ[0, 0] == (sc["occ"][v0] and sc["roi"][v0]) and val[0, 0] == sc["vals"][v0]
    assert occ.shape == (4, 1) and val.shape == (4, 1)


def test_contract_samples_known_answers_and_round_trip():
    """RaySamplerGPU.cuh:528-658: points inside |2x| <= 1 are copied; outside they land in the ball of radius 1 (|2x'| = 2 - 1/|2x|)"""
    o = np.array([[0.0, 0.0, 0.0], [0.1, -0.2, 0.3], [0.0, 0.0, 0.0]], F)
    se = np.array([[0, 3], [3, 3], [3, 7]], np.int32)  # ragged, with an empty ray
    p = np.array([[0.25, 0, 0], [1.0, 0, 0], [0, 4.0, 0], [0.5, 0, 0], [0, 0, -2.0], [3.0, 4.0, 0], [0.1, 0.1, 0.1]], F)
    z = np.arange(7, dtype=F).reshape(-1, 1)
    c, cz = S.contract_samples(o, se, p, z)
    # |2x| = 0.5 -> copy; |2x| = 2 -> factor 1.5, x' = 1.5 * 1 / 2 = 0.75; |2x| = 8 -> factor 1.875, y' = 1.875 * 4 / 8
    assert np.array_equal(c[0], p[0]) and cz[0, 0] == 0.0
    assert np.array_equal(c[1], np.array([0.75, 0, 0], F)) and cz[1, 0] == F(0.75)
    assert np.array_equal(c[2], np.array([0, 0.9375, 0], F)) and cz[2, 0] == F(0.9375)
    assert np.array_equal(c[3], p[3]) and cz[3, 0] == 3.0  # |2x| == 1 exactly: not contracted
    assert np.array_equal(c[4], np.array([0, 0, -0.875], F))  # |2x| = 4 -> factor 1.75, z' = 1.75 * -2 / 4
    assert np.array_equal(c[6], p[6]) and cz[6, 0] == 6.0
    assert (np.linalg.norm(c.astype(np.float64), axis=1) < 1.0).all()
    u, uz = S.contract_samples(o, se, c, cz, uncontract=True)
    assert np.allclose(u, p, rtol=2e-6, atol=0) and np.array_equal(u[[0, 3, 6]], p[[0, 3, 6]])
    # depths are distances from the ray's origin after the map
    assert np.allclose(uz[[1, 2, 4, 5], 0], np.linalg.norm(u[[1, 2, 4, 5]] - o[[0, 0, 2, 2]], axis=1), rtol=1e-6)
    e3, ez = S.contract_samples(o, np.zeros((3, 2), np.int32), np.zeros((0, 3), F), np.zeros((0, 1), F))
    assert e3.shape == (0, 3) and ez.shape == (0, 1)


def test_grid_points_and_density_maintenance_known_answers():
    """OccupancyGridGPU.cuh:31-218: voxel 0 sits at the lower-left corner of a grid centred at the origin; centres are half a voxel in;
    the centre of voxel i maps back to i; a single dense voxel occupies its 3x3x3 neighbourhood"""
    n, ext = 4, [1.0, 1.2, 0.9]
    idx = np.arange(n ** 3, dtype=np.int32)
    ll = S.grid_points(idx, n, ext, centre=False)
    c = S.grid_points(idx, n, ext, centre=True)
    assert np.array_equal(ll[0], np.array([-0.5, -0.6, -0.45], F)) and np.array_equal(ll[1], np.array([-0.25, -0.6, -0.45], F))
    assert np.array_equal(ll[2], np.array([-0.5, F(-0.25) * F(1.2), -0.45], F))  # Morton order: bit 1 is y
    assert np.array_equal(c[0], np.array([-0.375, F(-0.375) * F(1.2), F(-0.375) * F(0.9)], F))
    assert all(S.pos_to_lin_idx(c[i], n, ext) == i for i in range(n ** 3))
    j = S.grid_points(idx, n, ext, centre=True, jitter=True, rng=Pcg32())
    assert (np.abs(j - c) <= np.array(ext, F) / n / 2 + 1e-7).all() and np.abs(j - c).max() > 0.05
    assert all(S.pos_to_lin_idx(j[i], n, ext) == i for i in range(n ** 3))  # jitter stays inside the voxel
    g = np.zeros(n ** 3, F)
    g[S.morton3d(1, 1, 1)] = 1.0
    g2 = S.update_grid_values(idx[:8], np.full((8, 1), 0.25, F), 0.5, g)
    assert g2[S.morton3d(1, 1, 1)] == 0.5 and g2[0] == 0.25 and g2[8] == 0.0  # max(new, decay * old); untouched beyond the indices
    occ = S.update_grid_occupancy_density(idx, n, [1, 1, 1], 0.5, True, g, np.zeros(n ** 3, bool))
    assert occ.sum() == 27 and occ[S.morton3d(2, 2, 2)] and not occ[S.morton3d(3, 1, 1)]
    occ = S.update_grid_occupancy_density(idx, n, [1, 1, 1], 0.5, False, g, np.ones(n ** 3, bool))
    assert occ.sum() == 1 and occ[S.morton3d(1, 1, 1)]
    occ = S.update_grid_occupancy_density(idx[:4], n, [1, 1, 1], 0.5, False, g, np.ones(n ** 3, bool))
    assert occ.sum() == n ** 3 - 4  # only the listed voxels are rewritten


def test_sdf_occupancy_rule_known_answers():
    """OccupancyGridGPU.cuh:220-316: at the surface (|sdf| below half the voxel diagonal) the weight is beta / 4; far away it vanishes"""
    n = 8
    idx = np.arange(6, dtype=np.int32)
    g = np.zeros(n ** 3, F)
    g[:6] = [0.0, 0.05, -0.05, 0.5, -0.5, 0.2]
    beta = np.full((6, 1), 100.0, F)
    occ, w = S.update_grid_occupancy_sdf(idx, n, [1, 1, 1], beta, 1e-4, g, np.zeros(n ** 3, bool), return_weight=True)
    half_diag = np.sqrt(3.0) / 8 / 2  # 0.108
    assert w[0] == 25.0 and w[1] == 25.0 and w[2] == 25.0  # inside half a diagonal of the surface: d = 0 -> beta / 4
    want = 100.0 * np.exp(-100.0 * (0.2 - half_diag)) / (1 + np.exp(-100.0 * (0.2 - half_diag))) ** 2
    assert abs(w[5] - want) < 1e-4 * want and w[3] == w[4] and w[3] < 1e-12
    assert occ[:6].tolist() == [True, True, True, False, False, True] and not occ[6:].any()
