"""Parity against the REFERENCE'S OWN KERNELS: oracle/_ref/libvolsurfs_ref.so is the reference's
kernels/volsurfs/VolumeRenderingGPU.cuh compiled unmodified behind a C harness (oracle/ref_harness.cu, recipe oracle/build.py:build_ref).
Every test runs the reference kernel, the product (through the C ABI / the PyBridge-shaped shim) and the numpy restatement on the
same seeded inputs on the GPU box.  This pins both the oracle (oracle/compositing.py, oracle/importance.py) and the CUDA path to the
reference implementation itself.

Bars: integer outputs and single-product results bit-exact; sequential fp32 sums vs the product's shuffle scans within 1e-5
(re-association); oracle vs reference: bit-exact where no multiply-add chain exists (products, plain sums, cdf, median depth,
indices), within 1e-5 for the weighted sums (nvcc contracts a*b+c to FMA in the reference build, numpy does not; up to 1024 terms)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import ROOT, grad_err, rel_err
from oracle import compositing as oc
from oracle import importance as oi
from volsurfs_b200.synthetic import dense_layers, nerf_packets, pack_dense

pytestmark = pytest.mark.gpu
REF_PATH = ROOT / "oracle" / "_ref" / "libvolsurfs_ref.so"
TOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    if not REF_PATH.exists():
        pytest.skip("oracle/_ref/libvolsurfs_ref.so not built (python -m oracle.build where /root/reference is mounted)")
    lib = ctypes.CDLL(str(REF_PATH))
    assert lib.ref_abi_version() == 1
    return lib


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def _rsp(se):
    from volsurfs_b200.volsurfs import RaySamplesPacked

    rsp = RaySamplesPacked(0, 0, 0, 1)
    rsp.ray_start_end_idx = se.cuda()
    return rsp


def _cases():
    d = dense_layers(4096, 5, seed_offset=1)
    se, a = pack_dense(d["hit"], d["alpha"])
    yield "shells_k5", se, a.shape[0]
    p = nerf_packets(3000, seed_offset=3)
    yield "nerf_1024", p["se"], p["alpha"].shape[0]


def _ok(code):
    assert code == 0, f"reference harness returned CUDA error {code}"


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_forward_ops_vs_reference_kernels(ref, case):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    _, se, S = case
    N = se.shape[0]
    g = torch.Generator().manual_seed(S)
    x = (torch.rand(S, 1, generator=g) * 0.999 + 1e-3).cuda()
    v1, v3, w = torch.randn(S, 1, generator=g).cuda(), torch.randn(S, 3, generator=g).cuda(), torch.rand(S, 1, generator=g).cuda()
    rsp = _rsp(se)
    sed = rsp.ray_start_end_idx
    sen = se.numpy()

    T_r, bg_r = torch.zeros(S, 1, device="cuda"), torch.ones(N, 1, device="cuda")      # VolumeRendering.cu:45-47
    _ok(ref.ref_cumprod_fwd(P(sed), P(x), P(T_r), P(bg_r), N, S))
    T, bg = VR.cumprod_one_minus_alpha_to_transmittance(rsp, x)
    assert rel_err(T.cpu().numpy(), T_r.cpu().numpy(), floor=1e-30) < TOL and rel_err(bg.cpu().numpy(), bg_r.cpu().numpy(), floor=1e-30) < TOL
    To, bgo = oc.packed_cumprod_one_minus_alpha_to_transmittance(sen, x.cpu().numpy())
    assert np.array_equal(To, T_r.cpu().numpy()) and np.array_equal(bgo, bg_r.cpu().numpy())  # sequential products: oracle == reference

    for dim, v in ((1, v1), (3, v3)):
        out_r = torch.zeros(N, dim, device="cuda")
        _ok(ref.ref_integrate_fwd(P(sed), P(v), P(w), P(out_r), dim, N, S))
        out = VR.integrate_with_weights_1d(rsp, v, w) if dim == 1 else VR.integrate_with_weights_3d(rsp, v, w)
        assert grad_err(out.cpu().numpy(), out_r.cpu().numpy()) < TOL
        assert grad_err(oc.packed_integrate_with_weights(sen, v.cpu().numpy(), w.cpu().numpy()), out_r.cpu().numpy()) < TOL

    for d in (1, 2, 3, 32):
        v = torch.randn(S, d, generator=g).cuda()
        pr_r, ps_r = torch.zeros(N, d, device="cuda"), torch.zeros(S, d, device="cuda")
        _ok(ref.ref_sum_fwd(P(sed), P(v), P(pr_r), P(ps_r), d, N, S))
        pr, ps = VR.sum_over_rays(rsp, v)
        assert grad_err(pr.cpu().numpy(), pr_r.cpu().numpy()) < TOL and grad_err(ps.cpu().numpy(), ps_r.cpu().numpy()) < TOL
        pro, pso = oc.packed_sum_over_rays(sen, v.cpu().numpy())
        assert np.array_equal(pro, pr_r.cpu().numpy()) and np.array_equal(pso, ps_r.cpu().numpy())  # plain sums: no contraction possible

    for inverse in (0, 1):
        cs_r = torch.zeros(S, 1, device="cuda")
        _ok(ref.ref_cumsum(P(sed), P(v1), inverse, P(cs_r), N, S))
        cs = VR.cumsum_over_rays(rsp, v1, bool(inverse))
        assert grad_err(cs.cpu().numpy(), cs_r.cpu().numpy()) < TOL
        assert np.array_equal(oc.packed_cumsum_over_rays(sen, v1.cpu().numpy(), bool(inverse)), cs_r.cpu().numpy())


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_backward_ops_vs_reference_kernels(ref, case):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    _, se, S = case
    N = se.shape[0]
    g = torch.Generator().manual_seed(S + 1)
    x = torch.rand(S, 1, generator=g) * 0.999 + 1e-3
    x[::97] = 0.0  # the clamp_min(x, 1e-6) divisor (VolumeRenderingGPU.cuh:937)
    x = x.cuda()
    gT, gbg = torch.randn(S, 1, generator=g).cuda(), torch.randn(N, 1, generator=g).cuda()
    rsp = _rsp(se)
    sed = rsp.ray_start_end_idx
    sen = se.numpy()

    # reference path: forward kernel, python glue LV = gT*T (volume_rendering_funcs.py:151-160), reverse cumsum kernel, backward kernel
    T_r, bg_r = torch.zeros(S, 1, device="cuda"), torch.ones(N, 1, device="cuda")
    _ok(ref.ref_cumprod_fwd(P(sed), P(x), P(T_r), P(bg_r), N, S))
    LV = (gT * T_r).contiguous()
    cs_r = torch.zeros(S, 1, device="cuda")
    _ok(ref.ref_cumsum(P(sed), P(LV), 1, P(cs_r), N, S))
    dx_r = torch.zeros(S, 1, device="cuda")
    _ok(ref.ref_cumprod_bwd(P(sed), P(gT), P(gbg), P(x), P(T_r), P(bg_r), P(cs_r), P(dx_r), N, S))
    T, bg = VR.cumprod_one_minus_alpha_to_transmittance(rsp, x)
    cs = VR.cumsum_over_rays(rsp, gT * T, True)
    dx = VR.cumprod_one_minus_alpha_to_transmittance_backward(gT, gbg, rsp, x, T, bg, cs)
    dxf = VR.cumprod_backward_fused(gT, gbg, rsp, x, T, bg)
    scale = max(1.0, float(dx_r.abs().max()) * 1e-6)
    assert rel_err(dx.cpu().numpy(), dx_r.cpu().numpy(), floor=scale) < 5e-5
    assert rel_err(dxf.cpu().numpy(), dx_r.cpu().numpy(), floor=scale) < 5e-5
    dxo = oc.packed_cumprod_backward(sen, gT.cpu().numpy(), gbg.cpu().numpy(), x.cpu().numpy(), T_r.cpu().numpy(), bg_r.cpu().numpy(),
                                     cs_r.cpu().numpy())
    assert rel_err(dxo, dx_r.cpu().numpy(), floor=scale) < TOL

    w = torch.rand(S, 1, generator=g).cuda()
    for dim in (1, 3):
        v, go = torch.randn(S, dim, generator=g).cuda(), torch.randn(N, dim, generator=g).cuda()
        dv_r, dw_r = torch.zeros(S, dim, device="cuda"), torch.zeros(S, 1, device="cuda")
        res = torch.zeros(N, dim, device="cuda")
        _ok(ref.ref_integrate_bwd(P(sed), P(go), P(v), P(w), P(res), P(dv_r), P(dw_r), dim, N, S))
        VR.reference_bugs = True   # the reference reads channel [1] twice in the 3-D backward (VolumeRenderingGPU.cuh:1021)
        try:
            fn = VR.integrate_with_weights_1d_backward if dim == 1 else VR.integrate_with_weights_3d_backward
            dv, dw = fn(go, rsp, v, w, None)
        finally:
            VR.reference_bugs = False
        assert np.array_equal(dv.cpu().numpy(), dv_r.cpu().numpy())
        assert grad_err(dw.cpu().numpy(), dw_r.cpu().numpy()) < TOL           # g.v dot product: FMA contraction in the reference build
        dvo, dwo = oc.packed_integrate_with_weights_backward(sen, go.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy(), ref_bug=True)
        assert np.array_equal(dvo, dv_r.cpu().numpy()) and grad_err(dwo, dw_r.cpu().numpy()) < TOL
        if dim == 3:  # the product's default (mathematically correct) dw differs from the reference exactly by the z-channel term
            _, dw_ok = fn(go, rsp, v, w, None)
            ray_of = np.repeat(np.arange(N), np.maximum(sen[:, 1] - sen[:, 0], 0))
            fix = (go.cpu().numpy()[ray_of, 2] * (v.cpu().numpy()[:, 2] - v.cpu().numpy()[:, 1]))[:, None]
            assert grad_err(dw_ok.cpu().numpy(), dw_r.cpu().numpy() + fix) < 1e-5
    for dim in (1, 2, 3):
        v = torch.randn(S, dim, generator=g).cuda()
        gr, gs = torch.randn(N, dim, generator=g).cuda(), torch.randn(S, dim, generator=g).cuda()
        dv_r = torch.zeros(S, dim, device="cuda")
        _ok(ref.ref_sum_bwd(P(sed), P(gr), P(gs), P(v), P(dv_r), dim, N, S))
        dv = VR.sum_over_rays_backward(gr, gs, rsp, v)
        assert np.array_equal(dv.cpu().numpy(), dv_r.cpu().numpy())
        assert np.array_equal(oc.packed_sum_over_rays_backward(sen, gr.cpu().numpy(), gs.cpu().numpy(), v.cpu().numpy()), dv_r.cpu().numpy())


def _nerf_rsp(n_rays, seed, max_per_ray=96, mean=20.0):
    from volsurfs_b200.volsurfs import RaySamplesPacked

    # >= 2 samples per non-empty ray: a 1-sample segment makes the reference's binary search spin forever (VolumeRenderingGPU.cuh:481-505)
    p = nerf_packets(n_rays, seed_offset=seed, max_per_ray=max_per_ray, mean=mean, min_per_ray=2)
    S, N = p["alpha"].shape[0], p["se"].shape[0]
    assert int(((p["se"][:, 1] - p["se"][:, 0]) == 1).sum()) == 0
    rsp = RaySamplesPacked(N, S, 0, 1)
    rsp.ray_start_end_idx = p["se"].cuda()
    rsp.samples_z = p["z"].cuda()
    rsp.samples_dt = p["dt"].cuda()
    rsp.has_dt = True
    g = torch.Generator().manual_seed(seed)
    rsp.ray_o = torch.randn(N, 3, generator=g).cuda()
    rsp.ray_d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1).cuda()
    rsp.samples_dirs = rsp.ray_d[torch.repeat_interleave(torch.arange(N), (p["se"][:, 1] - p["se"][:, 0]).clamp(min=0).long())].contiguous()
    rsp.samples_3d = torch.randn(S, 3, generator=g).cuda()
    return rsp, p, g


def test_sdf2alpha_median_cdf_vs_reference_kernels(ref):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    rsp, p, g = _nerf_rsp(4000, 14)
    S, N = p["alpha"].shape[0], p["se"].shape[0]
    sed, sen = rsp.ray_start_end_idx, p["se"].numpy()
    sdf = ((torch.rand(S, 1, generator=g) - 0.5) * 0.2).cuda()
    beta = (torch.full((S, 1), 64.0) + torch.rand(S, 1, generator=g) * 200).cuda()
    a_r = torch.zeros(S, 1, device="cuda")
    _ok(ref.ref_sdf2alpha(P(sed), P(rsp.samples_dt), P(sdf), P(beta), P(a_r), N, S))
    a = VR.sdf2alpha(rsp, sdf, beta)
    # the reference evaluates sigmoid in double (1.0 / (1.0 + exp(-x)), VolumeRenderingGPU.cuh:181); alpha cancels when the two cdfs meet
    assert rel_err(a.cpu().numpy(), a_r.cpu().numpy(), floor=1e-2) < 1e-4
    assert rel_err(oc.packed_sdf2alpha(sen, p["dt"].numpy(), sdf.cpu().numpy(), beta.cpu().numpy()), a_r.cpu().numpy(), floor=1e-2) < 1e-4

    n = np.maximum(sen[:, 1] - sen[:, 0], 0)
    w = torch.rand(S, 1, generator=g)
    ray_of = torch.repeat_interleave(torch.arange(N), torch.from_numpy(n))
    w = (w / torch.zeros(N).index_add_(0, ray_of, w[:, 0])[ray_of].unsqueeze(1)).cuda()
    cdf_r = torch.zeros(S, 1, device="cuda")
    _ok(ref.ref_compute_cdf(P(sed), P(w), P(cdf_r), N, S))
    cdf = VR.compute_cdf(rsp, w)
    assert np.array_equal(cdf.cpu().numpy(), cdf_r.cpu().numpy())
    assert np.array_equal(oc.packed_compute_cdf(sen, w.cpu().numpy()), cdf_r.cpu().numpy())
    for thr in (0.5, 0.9, 2.0):
        md_r = torch.zeros(N, 1, device="cuda")
        _ok(ref.ref_median_depth(P(sed), P(rsp.samples_z), P(w), ctypes.c_float(thr), P(md_r), N, S))
        VR.reference_bugs = True   # the reference's fallback reads samples_z[nr_samples - 1] without idx_start (VolumeRenderingGPU.cuh:407)
        try:
            md = VR.median_depth_over_rays(rsp, w, thr)
        finally:
            VR.reference_bugs = False
        assert np.array_equal(md.cpu().numpy(), md_r.cpu().numpy()), thr
        assert np.array_equal(oc.packed_median_depth(sen, p["z"].numpy(), w.cpu().numpy(), thr, ref_bug=True), md_r.cpu().numpy())


def _ulp_close(a, b, ulps=4):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)))


@pytest.mark.parametrize("jitter", [0, 1])
def test_importance_sample_vs_reference_kernel(ref, jitter):
    from volsurfs_b200 import _lib
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    rsp, p, g = _nerf_rsp(3000, 21, max_per_ray=64, mean=24.0)
    S, N = p["alpha"].shape[0], p["se"].shape[0]
    sen = p["se"].numpy()
    # compute_cdf needs >= 2 samples per ray (VolumeRenderingGPU.cuh:443-447); 1-sample rays would hang the reference's binary search
    assert np.all((sen[:, 1] - sen[:, 0] != 1))
    w = torch.rand(S, 1, generator=g).cuda()
    cdf = VR.compute_cdf(rsp, w / 8.0)
    n_imp = 16
    state, inc = oi.PCG_DEFAULT_STATE, oi.PCG_DEFAULT_INC

    def fresh():
        return (torch.arange(S, S + N * n_imp, dtype=torch.int32, device="cuda").unsqueeze(1), torch.full((N * n_imp, 3), -1.0, device="cuda"),
                torch.full((N * n_imp, 3), -1.0, device="cuda"), torch.full((N * n_imp, 1), -1.0, device="cuda"),
                torch.full((N, 2), -1, dtype=torch.int32, device="cuda"))

    i_r, p_r, d_r, z_r, se_r = fresh()
    _ok(ref.ref_importance_sample(P(rsp.ray_o), P(rsp.ray_d), P(rsp.ray_start_end_idx), P(rsp.samples_z), P(cdf), N, S, n_imp,
                                  ctypes.c_uint64(state), ctypes.c_uint64(inc), jitter, P(i_r), P(p_r), P(d_r), P(z_r), P(se_r)))
    _, p_o, d_o, z_o, se_o = fresh()
    _lib.check(_lib.lib().vs_importance_sample(rsp.ray_o.data_ptr(), rsp.ray_d.data_ptr(), rsp.ray_start_end_idx.data_ptr(),
                                               rsp.samples_z.data_ptr(), cdf.data_ptr(), N, S, n_imp, state, inc, jitter, p_o.data_ptr(),
                                               d_o.data_ptr(), z_o.data_ptr(), se_o.data_ptr(), torch.cuda.current_stream().cuda_stream),
               "vs_importance_sample")
    torch.cuda.synchronize()
    assert torch.equal(se_o, se_r) and torch.equal(d_o, d_r)
    exact = float((z_o == z_r).float().mean())
    print(f"jitter={jitter}: z bit-exact on {exact:.4%} of the samples")
    assert _ulp_close(z_o.cpu().numpy(), z_r.cpu().numpy()) and _ulp_close(p_o.cpu().numpy(), p_r.cpu().numpy(), 8)
    # the numpy restatement against the reference kernel (first 300 rays; pure-python loops)
    m = 300
    o = oi.importance_sample(rsp.ray_o[:m].cpu().numpy(), rsp.ray_d[:m].cpu().numpy(), sen[:m], p["z"].numpy(), cdf.cpu().numpy(), n_imp,
                             bool(jitter))
    assert np.array_equal(o["ray_start_end_idx"], se_r[:m].cpu().numpy())
    assert np.allclose(o["samples_z"], z_r[:m * n_imp].cpu().numpy(), rtol=2e-6, atol=1e-7)

    # end to end through the shim: compacted packet, samples_idx numbered from S (VolumeRendering.cu:486-489)
    VR._rng_state, VR._rng_inc = state, inc
    imp = VR.importance_sample(rsp, cdf, n_imp, bool(jitter))
    keep = (se_r[:, 0] >= 0).cpu().numpy()
    rows = np.concatenate([np.arange(r * n_imp, (r + 1) * n_imp) for r in np.nonzero(keep)[0]])
    assert imp.get_max_nr_samples() == len(rows)
    assert np.array_equal(imp.samples_idx.cpu().numpy()[:, 0], S + rows)
    assert _ulp_close(imp.samples_z.cpu().numpy(), z_r.cpu().numpy()[rows])
    assert (VR._rng_state != state) == bool(jitter)  # the host generator moves on by 2^32 after a jittered launch
    VR._rng_state, VR._rng_inc = oi.PCG_DEFAULT_STATE, oi.PCG_DEFAULT_INC


@pytest.mark.parametrize("min_dist", [0.0, 0.01])
def test_combine_vs_reference_kernel(ref, min_dist):
    from oracle.packing import RaySamplesPackedNP  # noqa: F401  (compaction oracle lives there)
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    a, pa, g = _nerf_rsp(2500, 31, max_per_ray=64, mean=24.0)
    S1, N = pa["alpha"].shape[0], pa["se"].shape[0]
    w = torch.rand(S1, 1, generator=g).cuda()
    cdf = VR.compute_cdf(a, w / 8.0)
    b = VR.importance_sample(a, cdf, 8, False)   # second packet: same rays (rays without samples are empty in both)
    for pkt in (a, b):
        pkt.samples_values = torch.randn(pkt.get_max_nr_samples(), 1, generator=g).cuda()
        pkt.has_samples_values = True
    S2 = b.get_max_nr_samples()
    n1 = (a.ray_start_end_idx[:, 1] - a.ray_start_end_idx[:, 0]).clamp(min=0)
    n2 = (b.ray_start_end_idx[:, 1] - b.ray_start_end_idx[:, 0]).clamp(min=0)
    out_start = (torch.cumsum(n1 + n2, 0) - (n1 + n2)).to(torch.int32).contiguous()  # VolumeRendering.cu:595-603
    n = S1 + S2
    c_idx = torch.arange(n, dtype=torch.int32, device="cuda").unsqueeze(1)
    c_3d, c_d = torch.full((n, 3), -1.0, device="cuda"), torch.full((n, 3), -1.0, device="cuda")
    c_z, c_v = torch.full((n, 1), -1.0, device="cuda"), torch.full((n, 1), -1.0, device="cuda")
    c_se = torch.full((N, 2), -1, dtype=torch.int32, device="cuda")
    _ok(ref.ref_combine(N, ctypes.c_float(min_dist), 1, S1, P(a.ray_start_end_idx), P(a.samples_idx), P(a.samples_3d), P(a.samples_dirs),
                        P(a.samples_z), P(a.samples_values), S2, P(b.ray_start_end_idx), P(b.samples_idx), P(b.samples_3d), P(b.samples_dirs),
                        P(b.samples_z), P(b.samples_values), n, P(out_start), P(c_idx), P(c_3d), P(c_d), P(c_z), P(c_v), P(c_se)))
    comb = VR.combine_ray_samples_packets(a, b, min_dist)
    # compact the reference's uncompacted result on the host and compare everything bit for bit
    se_r = c_se.cpu().numpy()
    rows = np.concatenate([np.arange(s, e) for s, e in se_r if e > s])
    assert comb.get_max_nr_samples() == len(rows)
    for name, ref_t in (("samples_idx", c_idx), ("samples_3d", c_3d), ("samples_dirs", c_d), ("samples_z", c_z), ("samples_values", c_v)):
        assert np.array_equal(getattr(comb, name).cpu().numpy(), ref_t.cpu().numpy()[rows]), name
    cnt = np.maximum(se_r[:, 1] - se_r[:, 0], 0)
    start = np.cumsum(cnt) - cnt
    want_se = np.stack([start, start + cnt], 1).astype(np.int32)
    want_se[cnt == 0] = -1
    assert np.array_equal(comb.ray_start_end_idx.cpu().numpy(), want_se)
    assert not comb.has_dt and comb.has_samples_values and comb.is_compacted
    # the numpy restatement against the reference kernel (first 200 rays)
    m = 200
    cpu = lambda t: t.cpu().numpy()  # noqa: E731
    o = oi.combine_ray_samples_packets(cpu(a.ray_start_end_idx)[:m], cpu(a.samples_idx), cpu(a.samples_3d), cpu(a.samples_dirs), cpu(a.samples_z),
                                       cpu(a.samples_values), cpu(b.ray_start_end_idx)[:m], cpu(b.samples_idx), cpu(b.samples_3d),
                                       cpu(b.samples_dirs), cpu(b.samples_z), cpu(b.samples_values), min_dist)
    assert np.array_equal(o["ray_start_end_idx"], se_r[:m])
    last = int(se_r[:m, 1].max())
    rows_m = np.concatenate([np.arange(s, e) for s, e in se_r[:m] if e > s])
    assert np.array_equal(o["samples_z"][rows_m], cpu(c_z)[rows_m]) and np.array_equal(o["samples_idx"][rows_m], cpu(c_idx)[rows_m])
    assert last <= n
