"""GPU parity of the permutohedral hash encoding (csrc/permuto.cu, volsurfs_b200/encoding.py) through the C ABI against
(1) the REFERENCE'S OWN KERNELS — oracle/_ref/libpermuto_ref.so is submodules/permutohedral_encoding/kernels/.../EncodingGPU.cuh
    compiled unmodified behind a C harness (oracle/ref_permuto_harness.cu, recipe oracle/build.py:build_ref_permuto), and
(2) the numpy restatement oracle/permuto.py.

Bars: forward BIT-EXACT three ways — product == reference kernels == numpy restatement (the restatement emulates the a*b+c
contractions nvcc makes in the elevation, read from the SASS; at the fine levels the elevated coordinates are ~1e5 with an ulp of
0.008, so one differently rounded sum moves a barycentric weight by 1e-3: nothing short of the same roundings agrees there).
Position gradients bit-exact per level against the reference kernels; lattice gradients (unordered fp32 atomics on both sides, so
only the accumulation order differs) within 1e-5 under grad_err where slots receive few terms, and judged against the
fp64-accumulated restatement where thousands of mixed-sign terms meet in one slot (coarse levels)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import ROOT, grad_err
from oracle import permuto as op

pytestmark = pytest.mark.gpu
REF_PATH = ROOT / "oracle" / "_ref" / "libpermuto_ref.so"


@pytest.fixture(scope="module")
def ref():
    if not REF_PATH.exists():
        pytest.skip("oracle/_ref/libpermuto_ref.so not built (python -m oracle.build where /root/reference is mounted)")
    lib = ctypes.CDLL(str(REF_PATH))
    assert lib.ref_permuto_abi_version() == 1
    return lib


def P(t):
    return ctypes.c_void_p(None if t is None else t.data_ptr())


def _ok(code):
    assert code == 0, f"reference harness returned CUDA error {code}"


def _setup(n=20000, L=24, cap=1 << 18, seed=0, amp=1.0, pos_dim=3, fine=1e-4, concat=True):
    from volsurfs_b200.encoding import PermutoEncoding

    torch.manual_seed(seed)
    enc = PermutoEncoding(pos_dim, cap, L, 2, np.geomspace(1.0, fine, L), True, concat, 1.0)
    with torch.no_grad():
        enc.lattice_values.copy_(torch.randn(L, cap, 2) * amp)
    g = torch.Generator().manual_seed(seed + 1)
    pos = torch.rand(n, pos_dim, generator=g).cuda()
    return enc, pos


def _ref_forward(ref, enc, pos, window):
    n, L = pos.shape[0], enc.nr_levels
    sliced = torch.full((L + 2, 2, n), float("nan"), device="cuda")
    _ok(ref.ref_permuto_forward(P(pos), P(enc.lattice_values), P(enc.scale_factor), P(enc.random_shift_per_level), P(window), P(sliced), n,
                                enc.capacity, L, 1, ctypes.c_float(enc.concat_points_scaling)))
    return sliced


def _ref_backward(ref, enc, pos, window, grad_rows, want_pos=True):
    """grad_rows [n, 2*(L+2)] as autograd hands it over; returns the reference's (lattice grad [L,cap,2], positions grad [n,3])"""
    n, L = pos.shape[0], enc.nr_levels
    g_mono = grad_rows.reshape(n, L + 2, 2).permute(1, 2, 0).contiguous()       # funcs.py:51 (.contiguous() of the permuted view)
    d_lat = torch.zeros(L, 2, enc.capacity, device="cuda")
    d_pos = torch.zeros(3, n, device="cuda") if want_pos else None
    _ok(ref.ref_permuto_backward(P(pos), P(enc.lattice_values), P(enc.scale_factor), P(enc.random_shift_per_level), P(window), P(g_mono),
                                 P(d_lat), P(d_pos), n, enc.capacity, L, 1))
    return d_lat.permute(0, 2, 1).contiguous(), (d_pos.t().contiguous() if want_pos else None)    # Encoding.cu:202-203


def test_forward_bit_exact_vs_reference_kernels(ref):
    enc, pos = _setup()
    for t in (1.0, 0.47):
        window = enc.anneal_window if t == 1.0 else torch.from_numpy(op.cosine_easing_window(24, t * 24)).cuda()
        ours = enc(pos, window)
        theirs = _ref_forward(ref, enc, pos, window).permute(2, 0, 1).reshape(pos.shape[0], -1)       # modules.py:85
        assert ours.shape == theirs.shape == (pos.shape[0], 52)
        assert torch.equal(ours, theirs), f"max abs diff {(ours - theirs).abs().max().item():.3e}"


def test_forward_vs_restatement():
    enc, pos = _setup(n=6000)
    ours = enc(pos).detach().cpu().numpy()
    want = op.to_rows(op.forward(pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(),
                                 enc.random_shift_per_level.detach().cpu().numpy(), np.ones(24, np.float32), True, 1.0, fma=True))
    assert np.array_equal(ours, want), f"{(ours != want).sum()} entries differ, max {np.abs(ours - want).max():.3e}"


def test_restatement_vs_reference_kernels(ref):
    enc, pos = _setup(n=4000, seed=7)
    theirs = _ref_forward(ref, enc, pos, enc.anneal_window).cpu().numpy()
    want = op.forward(pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(),
                      enc.random_shift_per_level.detach().cpu().numpy(), np.ones(24, np.float32), True, 1.0, fma=True)
    assert np.array_equal(theirs, want), f"{(theirs != want).sum()} entries differ, max {np.abs(theirs - want).max():.3e}"


def test_backward_vs_reference_kernels_and_restatement(ref):
    enc, pos = _setup(n=30000, seed=3)
    n = pos.shape[0]
    g = torch.Generator().manual_seed(9)
    grad = torch.randn(n, 52, generator=g).cuda()
    window = torch.from_numpy(op.cosine_easing_window(24, 0.8 * 24)).cuda()
    d_lat, d_pos = enc._launch_backward(enc.lattice_values, pos, window, grad, None, None, want_lattice=True, want_positions=True)
    r_lat, r_pos = _ref_backward(ref, enc, pos, window, grad)
    # fp64-accumulated restatement
    o_lat, o_pos = op.backward(pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(),
                               enc.random_shift_per_level.detach().cpu().numpy(), window.cpu().numpy(), op.from_rows(grad.cpu().numpy()),
                               fma=True, dtype=np.float64)
    d_lat, r_lat = d_lat.cpu().numpy(), r_lat.cpu().numpy()
    # fine levels: a slot receives a handful of terms, the order of the atomics is immaterial
    assert grad_err(d_lat[12:], r_lat[12:]) < 1e-5
    assert grad_err(d_lat[12:], o_lat[12:]) < 1e-5
    # coarse levels: thousands of O(1) mixed-sign terms per slot in fp32, in an unspecified order on both sides — each side is
    # judged against fp64 (the product aggregates runs before the atomic, so it is usually the closer one)
    # Both errors are the maximum over ~10^5 slots of an order-dependent rounding sum, so they fluctuate from run to run by a factor
    # of a few (observed ours / reference between 0.4 and 3.3): the product is held to the absolute bar the reference kernels are
    # held to, and to the same order of magnitude as the reference's own error in that run.
    ours_err, ref_err = grad_err(d_lat, o_lat), grad_err(r_lat, o_lat)
    assert ours_err < 5e-3 and ours_err < max(1e-5, 8 * ref_err), (ours_err, ref_err)
    assert ref_err < 5e-3, ref_err
    # position gradient: per level bit-exact (one-hot windows), summed over levels in a different order (registers vs atomics)
    for lvl in (0, 7, 19):
        w1 = torch.zeros(24, device="cuda")
        w1[lvl] = 0.625
        _, a = enc._launch_backward(enc.lattice_values, pos, w1, grad, None, None, want_lattice=False, want_positions=True)
        _, b = _ref_backward(ref, enc, pos, w1, grad)
        assert torch.equal(a, b), lvl
    assert grad_err(d_pos.cpu().numpy(), r_pos.cpu().numpy()) < 1e-5
    assert grad_err(d_pos.cpu().numpy(), o_pos) < 1e-5


def test_coherent_positions_use_the_aggregated_atomics(ref):
    """camera-ray-like input: consecutive positions 1e-4 apart, so whole warps share simplex vertices on the coarse levels"""
    enc, _ = _setup(n=8)
    n = 50000
    t = torch.arange(n, dtype=torch.float32).cuda() * 1e-5
    pos = torch.stack([0.2 + t, 0.7 - 0.5 * t, 0.4 + 0.25 * t], dim=1).contiguous()
    grad = torch.randn(n, 52, generator=torch.Generator().manual_seed(1)).cuda()
    d_lat, _ = enc._launch_backward(enc.lattice_values, pos, enc.anneal_window, grad, None, None)
    r_lat, _ = _ref_backward(ref, enc, pos, enc.anneal_window, grad, want_pos=False)
    # coarse-level entries sum tens of thousands of mixed-sign terms in fp32 in either implementation: judge against fp64
    o_lat, _ = op.backward(pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(),
                           enc.random_shift_per_level.detach().cpu().numpy(), np.ones(24, np.float32), op.from_rows(grad.cpu().numpy()),
                           want_positions_grad=False, fma=True, dtype=np.float64)
    ours_err, ref_err = grad_err(d_lat.cpu().numpy(), o_lat), grad_err(r_lat.cpu().numpy(), o_lat)
    assert ours_err < 1e-5 or (ours_err <= 8 * ref_err and ours_err < 5e-3), (ours_err, ref_err)  # see the note in the test above


@pytest.mark.parametrize("pos_dim,cap,L", [(2, 1 << 12, 6), (4, 5003, 5), (3, 1000003, 4)])
def test_other_dims_and_capacities_vs_restatement(pos_dim, cap, L):
    enc, pos = _setup(n=3000, L=L, cap=cap, pos_dim=pos_dim, fine=1e-2, seed=pos_dim)
    ours = enc(pos).detach().cpu().numpy()
    args = (pos.cpu().numpy(), enc.lattice_values.detach().cpu().numpy(), enc.scale_factor.cpu().numpy(),
            enc.random_shift_per_level.detach().cpu().numpy(), np.ones(L, np.float32))
    want = op.to_rows(op.forward(*args, True, 1.0, fma=True))
    assert ours.shape == want.shape == (3000, 2 * (L + (pos_dim + 1) // 2))
    assert np.array_equal(ours, want), f"{(ours != want).sum()} entries differ, max {np.abs(ours - want).max():.3e}"
    grad = torch.randn(*ours.shape, generator=torch.Generator().manual_seed(2)).cuda()
    d_lat, d_pos = enc._launch_backward(enc.lattice_values, pos, enc.anneal_window, grad, None, None, want_lattice=True, want_positions=True)
    o_lat, o_pos = op.backward(*args, op.from_rows(grad.cpu().numpy()), fma=True, dtype=np.float64)
    # slots of the small tables receive tens of fp32 atomic terms in an order that changes from run to run: against the fp64 oracle
    # that is accumulation noise of ~1e-5 of the rms (observed 0.6e-5 .. 1.01e-5 over runs at cap = 5003), not a per-entry 1e-5
    assert grad_err(d_lat.cpu().numpy(), o_lat) < 3e-5
    assert grad_err(d_pos.cpu().numpy(), o_pos) < 1e-5


def test_permutohash_encoder_matches_the_reference_wrapper_semantics():
    """volsurfs_py/encodings/permutohash.py:68-96: bounding-box normalisation, out-of-bounds mask, remove_last_element"""
    from volsurfs_b200.encoding import PermutoHashEncoder

    torch.manual_seed(5)
    enc = PermutoHashEncoder(log2_hashmap_size=16, bb_sides=2.5)
    with torch.no_grad():
        enc.encoder.lattice_values.copy_(torch.randn_like(enc.encoder.lattice_values))
    pts = (torch.rand(5000, 3, generator=torch.Generator().manual_seed(6)) * 2.8 - 1.4).cuda()
    feats, oob = enc(pts)
    assert feats.shape == (5000, 51) and feats.is_contiguous() and oob.dtype == torch.bool
    bb = torch.tensor([2.5, 2.5, 2.5], device="cuda")
    want_oob = torch.logical_or((pts <= -bb / 2).any(dim=1), (pts >= bb / 2).any(dim=1))
    assert torch.equal(oob, want_oob) and 0 < int(oob.sum()) < 5000
    scaled = (pts * (1 / (bb / 2)) + 1) / 2                          # the reference's torch ops, in its order
    plain = enc.encoder(scaled, enc.window(None))
    assert torch.equal(feats, plain[:, :-1])
    # coarse-to-fine window
    enc.nr_iters_for_c2f = 1000
    f2, _ = enc(pts, iter_nr=300)
    w = enc.window(300)
    assert torch.allclose(f2[:, :48], feats[:, :48] * w.repeat_interleave(2)[None], rtol=1e-5, atol=1e-6)


def test_autograd_and_n_valid_gate():
    from volsurfs_b200.encoding import PermutoHashEncoder

    torch.manual_seed(8)
    enc = PermutoHashEncoder(log2_hashmap_size=14)
    with torch.no_grad():
        enc.encoder.lattice_values.copy_(torch.randn_like(enc.encoder.lattice_values))
    pts = (torch.rand(4096, 3, generator=torch.Generator().manual_seed(1)) * 1.8 - 0.9).cuda()
    upstream = torch.randn(4096, 51, generator=torch.Generator().manual_seed(2)).cuda()
    feats, _ = enc(pts)
    assert feats.requires_grad
    (feats * upstream).sum().backward()
    g_full = enc.encoder.lattice_values.grad.clone()
    # the same with only the first 1000 rows valid (count on the device, as in the fused pipeline)
    enc.encoder.lattice_values.grad = None
    n_valid = torch.tensor([1000], dtype=torch.int64, device="cuda")
    out = torch.full((4096, 51), -7.0, device="cuda")
    enc.encoder._launch_forward(enc.encoder.lattice_values, pts, enc.window(None), 51, enc.bb_sides, n_valid, out=out)
    assert torch.equal(out[:1000], feats[:1000].detach()) and bool((out[1000:] == -7.0).all())
    d_lat, _ = enc.encoder._launch_backward(enc.encoder.lattice_values, pts, enc.window(None), upstream, enc.bb_sides, n_valid)
    f1k, _ = enc(pts[:1000].contiguous())
    (f1k * upstream[:1000]).sum().backward()
    # two fp32 atomic accumulations of the same terms in different orders (4096-row vs 1000-row tiling; 2^14 slots collide heavily):
    # agreement to rounding noise of the accumulation, not to 1e-5 of each entry
    assert grad_err(d_lat.cpu().numpy(), enc.encoder.lattice_values.grad.cpu().numpy()) < 1e-4
    assert grad_err(g_full.cpu().numpy(), enc.encoder.lattice_values.grad.cpu().numpy()) > 1e-3     # and it is not the full gradient
    # position gradient flows (chain rule through the bounding-box map)
    pts_g = pts[:512].clone().requires_grad_(True)
    f, _ = enc(pts_g)
    (f * upstream[:512]).sum().backward()
    assert pts_g.grad is not None and bool(torch.isfinite(pts_g.grad).all()) and float(pts_g.grad.abs().max()) > 0


def test_full_size_adjoint_property():
    """size-independent property at the benchmark's scale (2^20 positions, 24 levels, 2^18 slots): <g, E(V)> == <E^T(g), V>"""
    enc, pos = _setup(n=1 << 20, seed=11)
    grad = torch.randn(1 << 20, 48, generator=torch.Generator().manual_seed(3)).cuda()
    out = enc(pos, out_cols=48).detach()
    d_lat, _ = enc._launch_backward(enc.lattice_values, pos, enc.anneal_window, grad, None, None)
    lhs = float((out.double() * grad.double()).sum())
    rhs = float((d_lat.double() * enc.lattice_values.detach().double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1.0, (lhs, rhs)


def test_lattice_backward_in_level_blocks_equals_the_full_backward():
    """``_launch_backward(levels=(l0, l1))`` (used to pipeline the gradient exchange with the backward, bench.py): four level blocks fill the
    same table as one full launch — equal up to the order of the fp32 atomics"""
    from volsurfs_b200.encoding import PermutoHashEncoder

    torch.manual_seed(3)
    enc = PermutoHashEncoder(log2_hashmap_size=16, bb_sides=2.0, device=torch.device("cuda", 0))
    e = enc.encoder
    with torch.no_grad():
        e.lattice_values.normal_(0.0, 0.1)
    n = 50000
    pos = (torch.rand(n, 3, device="cuda") - 0.5) * 1.2
    g = torch.randn(n, enc.output_dim, device="cuda")
    full, _ = e._launch_backward(e.lattice_values, pos, enc.window(None), g, enc.bb_sides, None, want_lattice=True)
    parts = torch.zeros_like(e.lattice_values)
    for b in range(4):
        e._launch_backward(e.lattice_values, pos, enc.window(None), g, enc.bb_sides, None, want_lattice=True, d_lattice=parts,
                           levels=(6 * b, 6 * b + 6))
    scale = float(full.abs().max())
    assert scale > 0 and float((parts - full).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("n,want_pos", [(20000, False), (12345, True), (77, False)])
def test_keyed_backward_matches_unkeyed(n, want_pos):
    """vs_permuto_backward_keyed with an ordering hint: the same contributions in another order.  Position gradients (no atomics: one
    thread owns a position) are bit-identical; lattice gradients agree to accumulation order (the bar of the unkeyed test)."""
    enc, pos = _setup(n=n, seed=5)
    # a packed-packet-like order: runs of 5 "layers" along a slowly moving point, plus ragged runs
    g = torch.Generator().manual_seed(9)
    key = (torch.arange(n) % 5).to(torch.int32)
    key[n // 2:] = torch.randint(0, 7, (n - n // 2,), generator=g, dtype=torch.int32)
    key = key.cuda()
    grad = torch.randn(n, enc.output_dims(), generator=torch.Generator().manual_seed(3)).cuda()
    w = enc.anneal_window
    dl0, dp0 = enc._launch_backward(enc.lattice_values, pos, w, grad, None, None, want_lattice=True, want_positions=want_pos)
    dl1, dp1 = enc._launch_backward(enc.lattice_values, pos, w, grad, None, None, want_lattice=True, want_positions=want_pos, order_key=key)
    torch.cuda.synchronize()
    # fine levels: few terms per slot; coarse levels: thousands of mixed-sign terms per slot, either order is ~1e-4 from the truth
    assert grad_err(dl1[12:].cpu().numpy(), dl0[12:].cpu().numpy()) < 1e-5
    assert grad_err(dl1.cpu().numpy(), dl0.cpu().numpy()) < 1e-3
    if want_pos:
        assert torch.equal(dp0, dp1)
