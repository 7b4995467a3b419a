"""CPU tests of the permutohedral-encoding restatement (oracle/permuto.py) and of the host-side mirror (volsurfs_b200/encoding.py).
The reference ships no vectors for this stage (its tests directory is empty); the pins are (1) a hand-worked lattice example,
(2) algebraic properties of the algorithm (partition of unity, simplex consistency, adjointness of forward and backward,
finite differences for the position gradient) and (3) on the GPU box the reference's own kernels (tests/test_gpu_permuto.py)."""
import math

import numpy as np
import pytest

from oracle import permuto as op

F = np.float32


def _encoder(L=8, cap=1 << 12, seed=3, coarse=1.0, fine=0.01, amp=1.0):
    enc = op.PermutoEncoding(3, cap, L, 2, np.geomspace(coarse, fine, L), True, True, 1.0, seed=seed)
    rng = np.random.default_rng(seed + 1)
    enc.lattice_values = (rng.standard_normal((L, cap, 2)) * amp).astype(F)
    return enc


def test_scale_factor_formula():
    s = op.scale_factors([1.0, 0.5], 3)
    for i in range(3):
        assert s[0, i] == F(1.0 / math.sqrt((i + 1) * (i + 2)))
        assert s[1, i] == F(1.0 / math.sqrt((i + 1) * (i + 2))) * F(2.0)


def test_worked_example_origin():
    """pos = 0, no shift: elevated = 0, closest 0-coloured point = origin, all differences tie -> rank = (0,1,2,3) from the strict '<'
    of EncodingGPU.cuh:156-163, barycentric = (1,0,0,0): the value is exactly the table entry of key (0,0,0) -> hash 0."""
    cap = 64
    lat = np.arange(cap * 2, dtype=F).reshape(1, cap, 2) + F(1)
    out = op.forward(np.zeros((1, 3), F), lat, op.scale_factors([1.0], 3), np.zeros((1, 3), F), np.ones(1, F), concat_points=False)
    assert out.shape == (1, 2, 1)
    assert out[0, 0, 0] == lat[0, 0, 0] and out[0, 1, 0] == lat[0, 0, 1]
    rem0, rank, bary = op._simplex(np.zeros((1, 3), F), np.zeros(3, F), op.scale_factors([1.0], 3)[0], False)
    assert rem0.tolist() == [[0, 0, 0, 0]] and rank.tolist() == [[0, 1, 2, 3]]
    assert bary[0, :4].tolist() == [1.0, 0.0, 0.0, 0.0]


def test_hash_is_the_reference_base_conversion():
    """hash(key) = ((k0*M + k1)*M + k2)*M mod 2^32, M = 2531011 (EncodingGPU.cuh:22-33), then mod capacity"""
    rem0 = np.array([[4, -8, 0, 4]], np.int32)
    rank = np.array([[0, 1, 2, 3]], np.int32)
    M = 2531011
    for cap in (1 << 18, 1000003):
        for r in range(4):
            key = [int(rem0[0, i]) + r - (4 if rank[0, i] > 3 - r else 0) for i in range(3)]
            k = 0
            for v in key:
                k = ((k + v) * M) & 0xFFFFFFFF
            assert op._vertex_index(rem0, rank, r, cap)[0] == k % cap


def test_simplex_invariants():
    enc = _encoder(L=12, fine=1e-4)
    pos = np.random.default_rng(0).random((4000, 3)).astype(F)
    for lvl in range(enc.nr_levels):
        rem0, rank, bary = op._simplex(pos, enc.random_shift_per_level[lvl], enc.scale[lvl], False)
        assert (np.sort(rank, axis=1) == np.arange(4)).all()             # ranks are a permutation
        assert (rem0.sum(axis=1) == 0).all() and (rem0 % 4 == rem0[:, :1] % 4).all()   # a lattice point of the hyperplane
        w = bary[:, :4]
        # fp32 cancellation: |elevated| reaches ~1e5 on the fine levels, so the weights carry ~|elevated| * 2^-24 of noise
        noise = 4e-7 * max(1.0, float(np.abs(enc.random_shift_per_level[lvl]).max() * enc.scale[lvl].max() * 3))
        assert np.abs(w.sum(axis=1) - 1).max() < 10 * noise + 1e-6
        assert w.min() > -(10 * noise + 1e-6)


def test_concat_points_and_row_layout():
    enc = _encoder(L=4)
    pos = np.random.default_rng(1).random((50, 3)).astype(F)
    rows = enc.forward(pos)
    assert rows.shape == (50, enc.output_dims()) and enc.output_dims() == 2 * (4 + 2)
    assert np.array_equal(rows[:, 8:11], pos) and (rows[:, 11] == 0).all()      # levels 4,5: (x,y), (z,0)
    mono = op.forward(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, enc.anneal_window, True, 1.0)
    assert np.array_equal(op.from_rows(rows), mono)


def test_window_scales_levels_linearly():
    enc = _encoder(L=6)
    pos = np.random.default_rng(2).random((200, 3)).astype(F)
    full = enc.forward(pos)
    win = op.cosine_easing_window(6, 0.45 * 6)
    part = enc.forward(pos, win)
    assert win[0] == 1 and win[-1] == 0 and 0 < win[2] < 1
    assert np.allclose(part[:, :12], full[:, :12] * np.repeat(win, 2)[None], rtol=1e-6, atol=1e-7)
    assert np.array_equal(part[:, 12:], full[:, 12:])


@pytest.mark.parametrize("fma", [False, True])
def test_backward_is_the_adjoint_of_forward(fma):
    enc = _encoder(L=10, fine=1e-3)
    rng = np.random.default_rng(4)
    pos = rng.random((3000, 3)).astype(F)
    g = rng.standard_normal((enc.nr_levels + 2, 2, 3000)).astype(F)
    win = op.cosine_easing_window(10, 7.3)
    sl = op.forward(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, win, fma=fma)
    g_lat, _ = op.backward(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, win, g, fma=fma, dtype=np.float64)
    lhs = float((g[:10].astype(np.float64) * sl[:10]).sum())
    rhs = float((g_lat * enc.lattice_values).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-3


def test_position_gradient_matches_finite_differences():
    enc = _encoder(L=5, coarse=1.0, fine=0.2)
    rng = np.random.default_rng(5)
    pos = rng.random((2000, 3)).astype(F)
    g = rng.standard_normal((7, 2, 2000)).astype(F)
    g[5:] = 0          # as in the reference, the concat-points levels pass no gradient
    w = np.ones(5, F)
    _, g_pos = op.backward(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, w, g, dtype=np.float64)
    eps = 2e-3
    for ax in range(3):
        pp, pm = pos.copy(), pos.copy()
        pp[:, ax] += eps
        pm[:, ax] -= eps
        fp = op.forward(pp, enc.lattice_values, enc.scale, enc.random_shift_per_level, w)
        fm = op.forward(pm, enc.lattice_values, enc.scale, enc.random_shift_per_level, w)
        fd = ((fp[:5].astype(np.float64) - fm[:5]) * g[:5]).sum(axis=(0, 1)) / (pp[:, ax].astype(np.float64) - pm[:, ax])
        err = np.abs(fd - g_pos[:, ax]) / np.maximum(np.abs(g_pos[:, ax]), 1.0)
        # piecewise linear: exact inside a simplex, off only for the few points whose +-eps stencil crosses a face
        assert np.median(err) < 5e-3 and (err < 5e-2).mean() > 0.9


def test_points_to_unit_cube():
    p = np.array([[0.0, 0.0, 0.0], [0.99, -0.5, 0.25], [1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], F)
    q, oob = op.volsurfs_points_to_unit_cube(p, 2.0)
    assert oob.tolist() == [False, False, True, True]
    assert np.array_equal(q[0], [0.5, 0.5, 0.5]) and q[1, 0] == (F(0.99) + F(1)) / F(2)


def test_host_mirror_shapes_without_gpu():
    import torch

    from volsurfs_b200 import encoding as ve

    w = ve.Coarse2Fine(24)(0.3)
    assert w.shape == (24,) and np.allclose(w.numpy(), op.cosine_easing_window(24, 0.3 * 24), atol=1e-6)
    assert torch.equal(ve.scale_factors(np.geomspace(1.0, 1e-4, 24), 3), torch.from_numpy(op.scale_factors(np.geomspace(1.0, 1e-4, 24), 3)))
    enc = ve.PermutoHashEncoder(log2_hashmap_size=10, device="cpu")
    assert enc.output_dim == 51 and enc.encoder.output_dims() == 52
    assert tuple(enc.encoder.lattice_values.shape) == (24, 1024, 2) and tuple(enc.encoder.random_shift_per_level.shape) == (24, 3)
    from volsurfs_b200 import _lib

    with pytest.raises(_lib.VolsurfsB200Error):       # no CPU fallback
        enc(torch.rand(8, 3))
    with pytest.raises(RuntimeError):
        ve.PermutoEncoding(3, 1024, 2, 4, [1.0, 0.5])


def test_c_twin_equals_the_numpy_oracle():
    """oracle/permuto_oracle.c (bench.py's multi-threaded CPU baseline of the encoder) against oracle/permuto.py with fma=False: forward
    rows bit for bit, lattice gradient bit for bit (same position-ordered accumulation)"""
    from oracle import permuto as op

    rng = np.random.default_rng(3)
    n, L, cap = 3000, 24, 2 ** 12
    enc = op.PermutoEncoding(3, cap, L, 2, np.geomspace(1.0, 1e-4, L), True, True, 1.0, seed=5)
    enc.lattice_values = rng.standard_normal(enc.lattice_values.shape).astype(np.float32)
    pos = rng.random((n, 3)).astype(np.float32)
    want = enc.forward(pos, fma=False)
    got = op.forward_rows_c(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, enc.anneal_window, True, 1.0)
    assert np.array_equal(got, want)
    g = rng.standard_normal(want.shape).astype(np.float32)
    want_lat, _ = op.backward(pos, enc.lattice_values, enc.scale, enc.random_shift_per_level, enc.anneal_window, op.from_rows(g), True, False,
                              fma=False)
    got_lat = op.backward_lattice_c(pos, enc.lattice_values.shape, enc.scale, enc.random_shift_per_level, enc.anneal_window, g)
    assert np.array_equal(got_lat, want_lat)
