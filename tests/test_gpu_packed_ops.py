"""GPU parity of the op-by-op packed VolumeRendering operators (drop-in API) against the numpy oracle that follows
kernels/volsurfs/VolumeRenderingGPU.cuh.  fp32, tolerance 1e-5 relative (sequential vs shuffle-scan association)."""
import numpy as np
import pytest
import torch

from conftest import grad_err, rel_err
from oracle import compositing as oc
from oracle.packing import RaySamplesPackedNP
from volsurfs_b200.synthetic import dense_layers, nerf_packets, pack_dense

pytestmark = pytest.mark.gpu
TOL = 1e-5


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
    p = nerf_packets(3000, seed_offset=5, max_per_ray=48, mean=12.0)
    yield "nerf_w16", p["se"], p["alpha"].shape[0]


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_forward_ops(case):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    _, se, S = case
    g = torch.Generator().manual_seed(S)
    x = torch.rand(S, 1, generator=g) * 0.999 + 1e-3
    v1 = torch.randn(S, 1, generator=g)
    v3 = torch.randn(S, 3, generator=g)
    w = torch.rand(S, 1, generator=g)
    rsp = _rsp(se)
    sen = se.numpy()

    T, bg = VR.cumprod_one_minus_alpha_to_transmittance(rsp, x.cuda())
    To, bgo = oc.packed_cumprod_one_minus_alpha_to_transmittance(sen, x.numpy())
    assert rel_err(T.cpu().numpy(), To, floor=1e-30) < TOL and rel_err(bg.cpu().numpy(), bgo, floor=1e-30) < TOL

    o1 = VR.integrate_with_weights_1d(rsp, v1.cuda(), w.cuda())
    o3 = VR.integrate_with_weights_3d(rsp, v3.cuda(), w.cuda())
    assert grad_err(o1.cpu().numpy(), oc.packed_integrate_with_weights(sen, v1.numpy(), w.numpy(), np.float64)) < TOL
    assert grad_err(o3.cpu().numpy(), oc.packed_integrate_with_weights(sen, v3.numpy(), w.numpy(), np.float64)) < TOL

    for d in (1, 2, 3, 32):
        v = torch.randn(S, d, generator=g)
        sr, ss = VR.sum_over_rays(rsp, v.cuda())
        sro, sso = oc.packed_sum_over_rays(sen, v.numpy(), np.float64)
        assert grad_err(sr.cpu().numpy(), sro) < TOL, d
        assert grad_err(ss.cpu().numpy(), sso) < TOL, d

    for inverse in (False, True):
        cs = VR.cumsum_over_rays(rsp, v1.cuda(), inverse)
        assert grad_err(cs.cpu().numpy(), oc.packed_cumsum_over_rays(sen, v1.numpy(), inverse, np.float64)) < TOL


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_backward_ops(case):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    _, se, S = case
    N = se.shape[0]
    g = torch.Generator().manual_seed(S + 1)
    x = torch.rand(S, 1, generator=g) * 0.999 + 1e-3
    x[::97] = 0.0  # exercise the clamp_min(x, 1e-6) divisor (VolumeRenderingGPU.cuh:937)
    gT = torch.randn(S, 1, generator=g)
    gbg = torch.randn(N, 1, generator=g)
    rsp = _rsp(se)
    sen = se.numpy()
    T, bg = VR.cumprod_one_minus_alpha_to_transmittance(rsp, x.cuda())
    # reference two-step path: python LV + cumsum_over_rays + backward kernel (volume_rendering_funcs.py:105-179)
    LV = gT.cuda() * T
    cs = VR.cumsum_over_rays(rsp, LV, True)
    dx = VR.cumprod_one_minus_alpha_to_transmittance_backward(gT.cuda(), gbg.cuda(), rsp, x.cuda(), T, bg, cs)
    dxf = VR.cumprod_backward_fused(gT.cuda(), gbg.cuda(), rsp, x.cuda(), T, bg)
    To, bgo = oc.packed_cumprod_one_minus_alpha_to_transmittance(sen, x.numpy(), np.float64)
    dxo = oc.packed_cumprod_backward_full(sen, gT.numpy(), gbg.numpy(), x.numpy(), To, bgo, np.float64)
    scale = max(1.0, float(np.abs(dxo).max()) * 1e-6)
    assert rel_err(dx.cpu().numpy(), dxo, floor=scale) < 5e-5
    assert rel_err(dxf.cpu().numpy(), dxo, floor=scale) < 5e-5

    w = torch.rand(S, 1, generator=g)
    for dim in (1, 3):
        v = torch.randn(S, dim, generator=g)
        go = torch.randn(N, dim, generator=g)
        for bug in (False, True):
            VR.reference_bugs = bug
            try:
                fn = VR.integrate_with_weights_1d_backward if dim == 1 else VR.integrate_with_weights_3d_backward
                dv, dw = fn(go.cuda(), rsp, v.cuda(), w.cuda(), None)
            finally:
                VR.reference_bugs = False
            dvo, dwo = oc.packed_integrate_with_weights_backward(sen, go.numpy(), v.numpy(), w.numpy(), ref_bug=bug)
            assert np.array_equal(dv.cpu().numpy(), dvo)          # single products: bit-exact
            assert np.array_equal(dw.cpu().numpy(), dwo), (dim, bug)  # same left-to-right order, no contraction
    for dim in (1, 2, 3):
        v = torch.randn(S, dim, generator=g)
        gr = torch.randn(N, dim, generator=g)
        gs = torch.randn(S, dim, generator=g)
        dv = VR.sum_over_rays_backward(gr.cuda(), gs.cuda(), rsp, v.cuda())
        assert np.array_equal(dv.cpu().numpy(), oc.packed_sum_over_rays_backward(sen, gr.numpy(), gs.numpy(), v.numpy()))


def test_reference_autograd_chain_vs_fused():
    """The four reference-named autograd Functions chained as in nerf.py:308-334 give the same image and gradients as
    CompositeFunc (bgT via 1 - sum w, nerf.py:323)."""
    from volsurfs_b200 import volume_rendering as vr

    p = nerf_packets(2000, seed_offset=8, max_per_ray=128, mean=30.0)
    rsp = _rsp(p["se"])
    rsp.samples_z = p["z"].cuda()
    alpha = p["alpha"].cuda().requires_grad_(True)
    rgb = p["rgb"].cuda().requires_grad_(True)
    m = vr.VolumeRenderingNeRF()
    T, _ = m.cumprod_one_minus_alpha_to_transmittance_module(rsp, 1 - alpha)
    w = alpha * T
    wsum, _ = m.sum_ray_module(rsp, w)
    pred = m.integrate_3d(rsp, rgb, w)
    loss = (pred * p["g_rgb"].cuda()).sum() + ((1 - wsum) * p["g_bgT"].cuda()).sum()
    loss.backward()
    a2 = p["alpha"].cuda().requires_grad_(True)
    c2 = p["rgb"].cuda().requires_grad_(True)
    rgb_f, depth_f, acc_f, bgT_f = vr.composite(rsp, a2, c2)
    loss2 = (rgb_f * p["g_rgb"].cuda()).sum() + (bgT_f * p["g_bgT"].cuda()).sum()
    loss2.backward()
    assert rel_err(pred.detach().cpu().numpy(), rgb_f.detach().cpu().numpy()) < TOL
    assert grad_err(c2.grad.cpu().numpy(), rgb.grad.cpu().numpy()) < TOL
    # alpha-gradients agree except where 1-alpha underflows the reference's clamp_min divisor
    ok = (1 - p["alpha"][:, 0]) > 1e-3
    assert rel_err(a2.grad.cpu().numpy()[ok], alpha.grad.cpu().numpy()[ok], floor=1e-1) < 1e-3


def test_update_dt_and_error_paths():
    from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR

    p = nerf_packets(500, seed_offset=9, max_per_ray=64, mean=10.0)
    S = p["alpha"].shape[0]
    rsp = RaySamplesPacked(500, S, 0, 1)
    rsp.ray_start_end_idx = p["se"].cuda()
    g = torch.Generator().manual_seed(3)
    z = torch.sort(torch.rand(S, 1, generator=g), 0).values
    rsp.samples_z = z.cuda()
    rsp.ray_exit = torch.full((500, 1), 1.2).cuda()
    rsp.ray_max_dt = torch.full((500, 1), 0.01).cuda()
    ref = RaySamplesPackedNP(500, S, 0, 1)
    ref.ray_start_end_idx = p["se"].numpy().copy()
    ref.samples_z = z.numpy().copy()
    ref.ray_exit[:] = 1.2
    ref.ray_max_dt[:] = 0.01
    for bg in (False, True):
        rsp.update_dt(bg)
        ref.update_dt(bg)
        assert np.array_equal(rsp.samples_dt.cpu().numpy(), ref.samples_dt)
    with pytest.raises(RuntimeError):
        VR.sum_over_rays(rsp, torch.zeros(S, 5).cuda())          # unsupported value dim (VolumeRendering.cu:243)
    with pytest.raises(RuntimeError):
        VR.integrate_with_weights_3d(rsp, torch.zeros(S, 2).cuda(), torch.zeros(S, 1).cuda())
    rsp.is_compacted = False
    with pytest.raises(RuntimeError):
        VR.cumsum_over_rays(rsp, torch.zeros(S, 1).cuda(), False)  # CHECK(is_compacted), VolumeRendering.cu:331
    with pytest.raises(ValueError):
        rsp.get_ray_max_dt(500)                                    # std::invalid_argument, RaySamplesPacked.cu:58-61


def test_next_ops_sdf2alpha_median_cdf():
    """the remaining simple VolumeRendering operators against the numpy oracle"""
    from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR

    p = nerf_packets(4000, seed_offset=14, max_per_ray=96, mean=20.0)
    S = p["alpha"].shape[0]
    N = p["se"].shape[0]
    sen = p["se"].numpy()
    g = torch.Generator().manual_seed(9)
    rsp = RaySamplesPacked(N, S, 0, 1)
    rsp.ray_start_end_idx = p["se"].cuda()
    rsp.samples_z = p["z"].cuda()
    rsp.samples_dt = p["dt"].cuda()
    rsp.has_dt = True
    sdf = (torch.rand(S, 1, generator=g) - 0.5) * 0.2
    beta = torch.full((S, 1), 64.0) + torch.rand(S, 1, generator=g) * 200
    alpha = VR.sdf2alpha(rsp, sdf.cuda(), beta.cuda())
    want = oc.packed_sdf2alpha(sen, p["dt"].numpy(), sdf.numpy(), beta.numpy())
    # expf (2 ulp) vs the oracle's double exp; alpha = (pc - nc + 1e-6)/(pc + 1e-6) cancels when pc ~ nc: absolute 1e-6 on [0,1]
    assert rel_err(alpha.cpu().numpy(), want, floor=1e-2) < 1e-4
    n = sen[:, 1] - sen[:, 0]
    last = sen[n > 0, 1] - 1
    assert np.all(alpha.cpu().numpy()[last] == 0)  # a ray's last sample keeps alpha 0
    # weights that sum to 1 per ray -> cdf snapping, and median depth
    w = torch.rand(S, 1, generator=g)
    ray_of = torch.repeat_interleave(torch.arange(N), torch.from_numpy(np.maximum(n, 0)))
    sums = torch.zeros(N).index_add_(0, ray_of, w[:, 0])
    w = w / sums[ray_of].unsqueeze(1)
    cdf = VR.compute_cdf(rsp, w.cuda())
    assert np.array_equal(cdf.cpu().numpy(), oc.packed_compute_cdf(sen, w.numpy()))
    for thr in (0.5, 0.9, 2.0):  # 2.0 is never reached: exercises the reference's fallback index
        for bug in (True, False):  # reference_bugs: the fallback reads samples_z[n-1] (VolumeRenderingGPU.cuh:407) / samples_z[start+n-1]
            VR.reference_bugs = bug
            try:
                md = VR.median_depth_over_rays(rsp, w.cuda(), thr)
            finally:
                VR.reference_bugs = False
            assert np.array_equal(md.cpu().numpy(), oc.packed_median_depth(sen, p["z"].numpy(), w.numpy(), thr, ref_bug=bug)), (thr, bug)
    rsp.has_dt = False
    with pytest.raises(RuntimeError):
        VR.sdf2alpha(rsp, sdf.cuda(), beta.cuda())
