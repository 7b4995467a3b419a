"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement (numpy, vectorised over positions, one level at a time) of the permutohedral-lattice hash encoding that feeds the
legacy appearance heads.  Citations are relative to /root/reference/submodules/permutohedral_encoding/:

* hash / modHash ................... kernels/permutohedral_encoding/EncodingGPU.cuh:22-45
* forward_gpu ...................... kernels/permutohedral_encoding/EncodingGPU.cuh:68-261 (host: src/Encoding.cu:55-113)
* backward_gpu (lattice values) .... kernels/permutohedral_encoding/EncodingGPU.cuh:264-416 (host: src/Encoding.cu:116-217)
* backward_gpu_only_pos ............ kernels/permutohedral_encoding/EncodingGPU.cuh:534-700
* scale factors .................... include/permutohedral_encoding/Encoding.cuh:53-67
* module glue (layouts, init) ...... src/pytorch_modules/modules.py:11-98, funcs.py:8-55, utils.py:5-17
* volsurfs wrapper ................. /root/reference/volsurfs_py/encodings/permutohash.py:10-99

Arithmetic: fp32 throughout, in the reference's evaluation order.  nvcc contracts ``a*b+c`` into FMAs in the reference build; pass
``fma=True`` to emulate that (the product of two fp32 values is exact in fp64, so ``fp32(fp64(a)*fp64(b)+fp64(c))`` is an FMA up to
rare double roundings).  Pin: tests/test_gpu_permuto.py runs the reference's own kernels (oracle/_ref/libpermuto_ref.so, compiled
from the reference sources) on the GPU box against this restatement and against the product; the reference ships no vectors.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
HASH_MULT = np.uint32(2531011)


def scale_factors(sigmas, pos_dim: int) -> np.ndarray:
    """Encoding.cuh:53-67: ``1/sqrt((i+1)(i+2))`` stored to an fp32 tensor element, then divided by the fp32 sigma.  torch divides a
    CUDA tensor by a host scalar as a multiplication with the fp32 reciprocal, so that is what is restated here."""
    sig = np.asarray(sigmas, dtype=np.float64).astype(F)       # pybind: std::vector<float>
    out = np.zeros((len(sig), pos_dim), F)
    for r in range(len(sig)):
        inv = F(1.0) / sig[r]
        for i in range(pos_dim):
            out[r, i] = F(1.0 / math.sqrt(float((i + 1) * (i + 2)))) * inv
    return out


def cosine_easing_window(num_freqs: int, alpha: float) -> np.ndarray:
    """utils.py:5-17"""
    x = np.clip(F(alpha) - np.arange(num_freqs, dtype=F), F(0), F(1))
    return (F(0.5) * (F(1) + np.cos(F(math.pi) * x + F(math.pi)))).astype(F)


def _mad(a, b, c, fma: bool):
    """a*b + c in fp32: separately rounded, or as one FMA"""
    if fma:
        return (a.astype(np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)
    return (a * b).astype(F) + c


def _simplex(pos, shift_l, scale_l, fma):
    """Elevation, rounding to the closest 0-coloured lattice point, ranks and barycentric weights of one level
    (EncodingGPU.cuh:128-205).  Returns rem0 [N,d+1] i32, rank [N,d+1] i32, bary [N,d+2] f32."""
    n, d = pos.shape
    elevated = np.zeros((n, d + 1), F)
    sm = np.zeros(n, F)
    for i in range(d, 0, -1):
        ps = (pos[:, i - 1] + shift_l[i - 1]).astype(F)
        cf = (ps * scale_l[i - 1]).astype(F)
        if not fma:
            elevated[:, i] = (sm - (F(i) * cf).astype(F)).astype(F)
            sm = (sm + cf).astype(F)
        elif i >= 3:
            # nvcc 12.9 (-O3, default -fmad=true), read from the SASS of forward_gpu / permuto_fwd_kernel for pos_dim 2, 3, 4:
            # cf is rounded once, i*cf is fused into the subtraction, the running sum adds the rounded cf
            elevated[:, i] = _mad(cf, F(-i), sm, True)
            sm = (sm + cf).astype(F)
        elif i == 2:
            # 2*cf is strength-reduced to cf + cf and ONE of the two products is fused: fma(ps, scale, cf)
            two_cf = _mad(ps, scale_l[i - 1], cf, True)
            elevated[:, i] = (sm - two_cf).astype(F)
            sm = (sm + cf).astype(F)
        else:
            # i == 1: cf is never materialised; both uses fuse the product (fma(-ps, scale, sm) and fma(ps, scale, sm))
            elevated[:, 1] = _mad(-ps, scale_l[0], sm, True)
            sm = _mad(ps, scale_l[0], sm, True)
    elevated[:, 0] = sm

    inv = F(1.0) / F(d + 1)                                    # the literal (1.0f / (pos_dim + 1))
    v = (elevated * inv).astype(F)
    up = (np.ceil(v) * F(d + 1)).astype(F)
    down = (np.floor(v) * F(d + 1)).astype(F)
    take_up = (up - elevated).astype(F) < (elevated - down).astype(F)
    rem0 = np.where(take_up, up, down).astype(np.int32)
    s = rem0.sum(axis=1, dtype=np.int64)
    s = np.fix(s / (d + 1)).astype(np.int32)                   # C integer division

    diff = (elevated - rem0.astype(F)).astype(F)
    rank = np.zeros((n, d + 1), np.int32)
    for i in range(d):
        for j in range(i + 1, d + 1):
            lt = diff[:, i] < diff[:, j]
            rank[:, i] += lt
            rank[:, j] += ~lt
    rank += s[:, None]
    lo, hi = rank < 0, rank > d
    rank = np.where(lo, rank + (d + 1), np.where(hi, rank - (d + 1), rank)).astype(np.int32)
    rem0 = np.where(lo, rem0 + (d + 1), np.where(hi, rem0 - (d + 1), rem0)).astype(np.int32)

    bary = np.zeros((n, d + 2), F)
    rows = np.arange(n)
    for i in range(d + 1):
        delta = ((elevated[:, i] - rem0[:, i].astype(F)).astype(F) * inv).astype(F)
        bary[rows, d - rank[:, i]] += delta
        bary[rows, d + 1 - rank[:, i]] -= delta
    bary[:, 0] = bary[:, 0] + (F(1.0) + bary[:, d + 1]).astype(F)
    return rem0, rank, bary


def _vertex_index(rem0, rank, remainder, capacity):
    """key of the simplex vertex with this remainder and its hash slot (EncodingGPU.cuh:22-45,216-227)"""
    n, d1 = rem0.shape
    d = d1 - 1
    k = np.zeros(n, np.uint32)
    with np.errstate(over="ignore"):
        for i in range(d):
            key = rem0[:, i] + remainder
            key = np.where(rank[:, i] > d - remainder, key - (d + 1), key).astype(np.int32)
            k = (k + key.astype(np.uint32)).astype(np.uint32)
            k = (k * HASH_MULT).astype(np.uint32)
    return (k % np.uint32(capacity)).astype(np.int64)


def n_extra_levels(pos_dim: int, val_dim: int, concat_points: bool) -> int:
    return int(math.ceil(float(pos_dim) / val_dim)) if concat_points else 0      # Encoding.cu:72-75


def forward(positions, lattice_values, scale, shift, window, concat_points=True, points_scaling=1.0, fma=False):
    """positions [N,d] f32, lattice_values [L,capacity,2] f32, scale/shift [L,d] f32, window [L] f32.
    Returns sliced values in the reference's monolithic layout [L+extra, 2, N] (Encoding.cu:79)."""
    pos = np.ascontiguousarray(positions, F)
    n, d = pos.shape
    L, cap, vd = lattice_values.shape
    assert vd == 2
    extra = n_extra_levels(d, vd, concat_points)
    out = np.zeros((L + extra, vd, n), F)
    for lvl in range(L):
        rem0, rank, bary = _simplex(pos, shift[lvl], scale[lvl], fma)
        acc = np.zeros((n, 2), F)
        for r in range(d + 1):
            idx = _vertex_index(rem0, rank, r, cap)
            w = (bary[:, r] * F(window[lvl])).astype(F)
            val = lattice_values[lvl, idx]
            acc[:, 0] = _mad(val[:, 0], w, acc[:, 0], fma)
            acc[:, 1] = _mad(val[:, 1], w, acc[:, 1], fma)
        out[lvl, 0], out[lvl, 1] = acc[:, 0], acc[:, 1]
    for e in range(extra):                                      # EncodingGPU.cuh:108-124
        for i in range(vd):
            src = i + e * vd
            out[L + e, i] = (pos[:, src] * F(points_scaling)).astype(F) if src < d else F(0)
    return out


def to_rows(sliced):
    """modules.py:85: [levels, val, N] -> [N, levels*val]"""
    return np.ascontiguousarray(np.transpose(sliced, (2, 0, 1)).reshape(sliced.shape[2], -1))


def from_rows(rows, val_dim=2):
    """inverse of to_rows (the layout autograd hands to lattice.backward)"""
    n, c = rows.shape
    return np.ascontiguousarray(rows.reshape(n, c // val_dim, val_dim).transpose(1, 2, 0))


def backward(positions, lattice_values, scale, shift, window, grad_sliced, concat_points=True, want_positions_grad=True, fma=False,
             dtype=np.float32):
    """grad_sliced [L+extra,2,N].  Returns (lattice_values_grad [L,capacity,2], positions_grad [N,d] or None) — the permuted /
    transposed views the reference returns (Encoding.cu:202-203).  Accumulation runs in position order in ``dtype`` (the reference
    accumulates with unordered fp32 atomics).  As in the reference, the concat-points levels pass no gradient to the positions."""
    pos = np.ascontiguousarray(positions, F)
    n, d = pos.shape
    L, cap, vd = lattice_values.shape
    g_lat = np.zeros((L, cap, vd), dtype)
    g_pos = np.zeros((n, d), dtype) if want_positions_grad else None
    inv = F(1.0) / F(d + 1)
    rows = np.arange(n)
    for lvl in range(L):
        rem0, rank, bary = _simplex(pos, shift[lvl], scale[lvl], fma)
        gx, gy = np.asarray(grad_sliced[lvl, 0], F), np.asarray(grad_sliced[lvl, 1], F)
        wl = F(window[lvl])
        dl_db = np.zeros((n, d + 2), F)
        for r in range(d + 1):
            idx = _vertex_index(rem0, rank, r, cap)
            w = (bary[:, r] * wl).astype(F)
            np.add.at(g_lat[lvl, :, 0], idx, (gx * w).astype(F).astype(dtype))
            np.add.at(g_lat[lvl, :, 1], idx, (gy * w).astype(F).astype(dtype))
            if want_positions_grad:
                val = lattice_values[lvl, idx]
                t = _mad((val[:, 0] * wl).astype(F), gx, dl_db[:, r], fma)
                dl_db[:, r] = _mad((val[:, 1] * wl).astype(F), gy, t, fma)
        if not want_positions_grad:
            continue
        dl_db[:, d + 1] += dl_db[:, 0]
        dl_de = np.zeros((n, d + 1), F)
        for i in range(d + 1):
            dl_de[:, i] = _mad(dl_db[rows, d - rank[:, i]], inv, dl_de[:, i], fma)
            dl_de[:, i] = _mad(dl_db[rows, d + 1 - rank[:, i]], -inv, dl_de[:, i], fma)
        for i in range(d):
            acc = np.zeros(n, F)
            for j in range(i + 1):
                acc = _mad(dl_de[:, j], scale[lvl, i], acc, fma)
            acc = _mad((dl_de[:, i + 1] * scale[lvl, i]).astype(F), F(-(i + 1)), acc, fma)
            g_pos[:, i] += acc.astype(dtype)
    return g_lat, g_pos


class PermutoEncoding:
    """numpy twin of permutohedral_encoding.PermutoEncoding (modules.py:11-98) for the tests: same constructor arguments, same
    parameter shapes; forward returns [N, output_dims()]."""

    def __init__(self, pos_dim, capacity, nr_levels, nr_feat_per_level, scale_per_level, appply_random_shift_per_level=True,
                 concat_points=False, concat_points_scaling=1.0, seed=0):
        assert nr_feat_per_level == 2                            # Encoding.cuh:153-156
        rng = np.random.default_rng(seed)
        self.pos_dim, self.capacity, self.nr_levels, self.nr_feat_per_level = pos_dim, capacity, nr_levels, nr_feat_per_level
        self.concat_points, self.concat_points_scaling = concat_points, concat_points_scaling
        self.scale = scale_factors(scale_per_level, pos_dim)
        self.lattice_values = (rng.standard_normal((nr_levels, capacity, 2)) * 1e-5).astype(F)
        self.random_shift_per_level = ((rng.standard_normal((nr_levels, pos_dim)) * 10).astype(F) if appply_random_shift_per_level
                                       else np.zeros((nr_levels, pos_dim), F))
        self.anneal_window = np.ones(nr_levels, F)

    def output_dims(self):
        return self.nr_feat_per_level * (self.nr_levels + n_extra_levels(self.pos_dim, self.nr_feat_per_level, self.concat_points))

    def forward(self, positions, anneal_window=None, fma=False):
        w = self.anneal_window if anneal_window is None else np.asarray(anneal_window, F)
        return to_rows(forward(positions, self.lattice_values, self.scale, self.random_shift_per_level, w, self.concat_points,
                               self.concat_points_scaling, fma=fma))


def volsurfs_points_to_unit_cube(points, bb_sides=2.0):
    """permutohash.py:77-86: out-of-bounds mask and the affine map of the bounding box onto [0,1]"""
    p = np.asarray(points, F)
    half = F(bb_sides) / F(2)
    oob = np.logical_or((p <= -half).any(axis=1), (p >= half).any(axis=1))
    scaled = (p * (F(1) / half)).astype(F)
    return ((scaled + F(1)) / F(2)).astype(F), oob


# ---- the C twin (oracle/permuto_oracle.c): same arithmetic with fma=False, all host threads; used by bench.py's CPU baseline --------------
_clib = None


def _c():
    global _clib
    if _clib is None:
        import ctypes

        from . import build as _build

        lib = ctypes.CDLL(str(_build.build_permuto_c()))
        P, I64, I, Fl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
        lib.vpo_forward.argtypes = [P, I64, I, P, I, I64, P, P, P, I, Fl, P]
        lib.vpo_backward_lattice.argtypes = [P, I64, I, I, I64, P, P, P, P, I, P]
        _clib = lib
    return _clib


def forward_rows_c(positions, lattice_values, scale, shift, window, concat_points=True, points_scaling=1.0):
    """``to_rows(forward(...))`` computed by the C twin: [N, 2*(L+extra)]"""
    pos = np.ascontiguousarray(positions, F)
    lat = np.ascontiguousarray(lattice_values, F)
    sc, sh, w = (np.ascontiguousarray(x, F) for x in (scale, shift, window))
    n, d = pos.shape
    L, cap, _ = lat.shape
    out = np.empty((n, 2 * (L + n_extra_levels(d, 2, concat_points))), F)
    _c().vpo_forward(pos.ctypes.data, n, d, lat.ctypes.data, L, cap, sc.ctypes.data, sh.ctypes.data, w.ctypes.data, int(concat_points),
                     float(points_scaling), out.ctypes.data)
    return out


def backward_lattice_c(positions, lattice_shape, scale, shift, window, grad_rows, out=None):
    """lattice gradient [L,capacity,2] from the gradient of the rows ([N, 2*(L+extra)]), C twin; ``out``: accumulate into this table"""
    pos = np.ascontiguousarray(positions, F)
    g = np.ascontiguousarray(grad_rows, F)
    sc, sh, w = (np.ascontiguousarray(x, F) for x in (scale, shift, window))
    n, d = pos.shape
    L, cap, _ = lattice_shape
    g_lat = np.zeros((L, cap, 2), F) if out is None else out
    _c().vpo_backward_lattice(pos.ctypes.data, n, d, L, cap, sc.ctypes.data, sh.ctypes.data, w.ctypes.data, g.ctypes.data, g.shape[1],
                              g_lat.ctypes.data)
    return g_lat
