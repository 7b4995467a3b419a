"""GPU parity of the fused compositor (C ABI vs_composite_fwd/bwd through the python shim) against the oracle.
Tolerance: 1e-5 relative (BASELINE.json north_star), rel = |a-b| / max(|b|, floor)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, grad_err, rel_err
from oracle import compositing as oc
from volsurfs_b200.synthetic import all_hit_packed, composite_bytes, dense_layers, nerf_packets, pack_dense

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rsp(se):
    from volsurfs_b200.volsurfs import RaySamplesPacked

    rsp = RaySamplesPacked(0, 0, 0, 1)
    rsp.ray_start_end_idx = se.cuda()
    return rsp


def _run(se, alpha, rgb, z, grads, mode, need_dz=True):
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    rsp = _rsp(se)
    a, c, zz = alpha.cuda(), rgb.cuda(), z.cuda()
    fwd = VR.composite(rsp, a, c, zz, return_weights=True, mode=mode)
    bwd = VR.composite_backward(rsp, a, c, zz, grads["g_rgb"].cuda(), grads["g_depth"].cuda(), grads["g_acc"].cuda(),
                                grads["g_bgT"].cuda(), need_dz=need_dz, mode=mode)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in fwd], [None if t is None else t.cpu().numpy() for t in bwd]


def _check(se, alpha, rgb, z, grads, mode):
    fwd, bwd = _run(se, alpha, rgb, z, grads, mode)
    sen = se.numpy()
    o = oc.fused_composite_forward(sen, alpha.numpy(), rgb.numpy(), z.numpy(), dtype=np.float64)
    for got, key in zip(fwd, ("rgb", "depth", "acc", "bgT", "weights", "transmittance")):
        assert rel_err(got, o[key], floor=1e-6) < TOL, key
    ob = oc.fused_composite_backward(sen, alpha.numpy(), rgb.numpy(), z.numpy(), grads["g_rgb"].numpy(), grads["g_depth"].numpy(),
                                     grads["g_acc"].numpy(), grads["g_bgT"].numpy())
    for got, key in zip(bwd, ("d_alpha", "d_rgb", "d_z")):
        assert grad_err(got, ob[key]) < TOL, key


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_c1_shells_vs_oracle(mode):
    """BASELINE config 1: 4096 rays x 5 layers, Bernoulli(0.8) hits, exact 0/1 alphas, empty rays."""
    d = dense_layers(4096, 5, seed_offset=1)
    se, a, c, z = pack_dense(d["hit"], d["alpha"], d["rgb"], d["z"])
    assert (se[:, 0] == -1).any(), "config must contain empty rays"
    _check(se, a, c, z, d, mode)


@pytest.mark.parametrize("mode", [0, 2])
def test_golden_reference_lines(mode):
    """CUDA result vs the vectors produced by the reference's own dense torch lines (fp32 variant)."""
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    for name in ("dense_composite_k5", "dense_composite_k1", "dense_composite_k9"):
        g = np.load(GOLDEN / f"{name}.npz")
        hit = torch.from_numpy(g["hit"])
        se, a, c, z = pack_dense(hit, torch.from_numpy(g["alpha"]), torch.from_numpy(g["rgb"]), torch.from_numpy(g["z"]))
        N = hit.shape[0]
        grads = {"g_rgb": torch.from_numpy(g["g_rgb"]), "g_depth": torch.zeros(N, 1), "g_acc": torch.zeros(N, 1),
                 "g_bgT": torch.from_numpy(g["g_bgT"])}
        fwd, bwd = _run(se, a, c, z, grads, mode, need_dz=False)
        assert rel_err(fwd[0], g["fp32_rgb_fg"]) < TOL
        assert rel_err(fwd[3], g["fp32_bg_transmittance"]) < TOL
        ray, j = np.nonzero(g["hit"][:, ::-1])
        lay = hit.shape[1] - 1 - j
        assert rel_err(fwd[4][:, 0], g["fp32_weights"][ray, lay, 0]) < TOL
        assert grad_err(bwd[0][:, 0], g["fp32_d_alpha"][ray, lay, 0]) < TOL
        assert grad_err(bwd[1], g["fp32_d_rgb"][ray, lay]) < TOL


@pytest.mark.parametrize("mode", [0, 2, 3, 5, 6, 7, 8])
def test_c3_nerf_packets_vs_oracle(mode):
    """Variable-length packets up to 1024 samples/ray with 35 % empty rays (config 3 shape, 6k-ray subset)."""
    p = nerf_packets(6000, seed_offset=3)
    assert int(p["counts"].max()) == 1024 and int((p["counts"] == 0).sum()) > 1000
    _check(p["se"], p["alpha"], p["rgb"], p["z"], p, mode)


def test_long_rays_spill_path():
    """rays longer than W*W = 1024 samples exercise the scratch path of the backward scan kernel"""
    p = nerf_packets(40, seed_offset=4, max_per_ray=3000, mean=1500.0, sigma=0.4, p_empty=0.1)
    assert int(p["counts"].max()) > 1024
    for mode in (2, 3, 5, 8):  # 8: ring family (spills beyond 32 chunks of 128); 2: one sample per lane (spills beyond 32*32), 5: quad per lane with 8 lanes (spills beyond 4*8*8)
        _check(p["se"], p["alpha"], p["rgb"], p["z"], p, mode)


def _packets_from_counts(counts, seed):
    g = torch.Generator().manual_seed(seed)
    cnt = torch.tensor(counts, dtype=torch.int64)
    n, S = cnt.numel(), int(cnt.sum())
    start = torch.cumsum(cnt, 0) - cnt
    se = torch.stack([start, start + cnt], 1).to(torch.int32)
    se[cnt == 0] = -1
    alpha = (torch.rand(S, 1, generator=g) * (torch.rand(S, 1, generator=g) > 0.5) * 0.2).contiguous()
    return {"se": se.contiguous(), "alpha": alpha, "rgb": torch.rand(S, 3, generator=g), "z": torch.rand(S, 1, generator=g) + 0.5,
            "g_rgb": torch.randn(n, 3, generator=g), "g_depth": torch.randn(n, 1, generator=g), "g_acc": torch.randn(n, 1, generator=g),
            "g_bgT": torch.randn(n, 1, generator=g)}


@pytest.mark.parametrize("counts", [
    [5] * 1000 + [700] * 10 + [0] * 7 + [3] * 50,       # mean 12: 8-lane groups, the 700-sample rays spill (more than 8 chunks of 32)
    [90] * 300 + [0] * 100 + [1500] * 4 + [1] * 33,     # mean 76: 16-lane groups, 1500-sample rays spill (more than 16 chunks of 64)
    [5000] * 3 + [300] * 40 + [0] * 5,                  # mean 562: 32-lane groups, 5000-sample rays spill (more than 32 chunks of 128)
], ids=["w8", "w16", "w32"])
def test_ring_family_group_widths_and_spill(counts):
    p = _packets_from_counts(counts, seed=len(counts))
    _check(p["se"], p["alpha"], p["rgb"], p["z"], p, 8)


@pytest.mark.parametrize("n_rays,K", [(1, 5), (255, 5), (257, 9), (1000, 1)])
def test_edge_sizes(n_rays, K):
    d = all_hit_packed(n_rays, K)
    for mode in (0, 1, 2, 3, 4, 5, 7, 8):
        _check(d["se"], d["alpha"], d["rgb"], d["z"], d, mode)


def test_all_empty_and_zero_rays():
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    se = torch.full((300, 2), -1, dtype=torch.int32)
    e1, e3 = torch.zeros(0, 1), torch.zeros(0, 3)
    for mode in (0, 1, 2, 3, 4, 6, 8):
        rgb, depth, acc, bgT = VR.composite(_rsp(se), e1.cuda(), e3.cuda(), e1.cuda(), mode=mode)
        assert torch.all(rgb == 0) and torch.all(depth == 0) and torch.all(acc == 0) and torch.all(bgT == 1)
    rgb, depth, acc, bgT = VR.composite(_rsp(torch.zeros((0, 2), dtype=torch.int32)), e1.cuda(), e3.cuda(), e1.cuda())
    assert rgb.shape == (0, 3) and bgT.shape == (0, 1)


def test_tile_fallback_on_skewed_and_gapped_layouts():
    """tile kernels must fall back per tile when a tile's sample range exceeds the shared-memory capacity
    (skewed counts) or is not contiguous (uncompacted slot layout with gaps)."""
    g = torch.Generator().manual_seed(5)
    n = 2048
    cnt = torch.ones(n, dtype=torch.int64)
    cnt[100:140] = 300  # one heavy tile, mean stays < 8
    start = torch.cumsum(cnt, 0) - cnt
    se = torch.stack([start, start + cnt], 1).to(torch.int32)
    S = int(cnt.sum())
    d = {"alpha": torch.rand(S, 1, generator=g), "rgb": torch.rand(S, 3, generator=g), "z": torch.rand(S, 1, generator=g),
         "g_rgb": torch.randn(n, 3, generator=g), "g_depth": torch.randn(n, 1, generator=g), "g_acc": torch.randn(n, 1, generator=g),
         "g_bgT": torch.randn(n, 1, generator=g)}
    _check(se, d["alpha"], d["rgb"], d["z"], d, 1)
    _check(se, d["alpha"], d["rgb"], d["z"], d, 4)
    # gapped: ray r owns slots [r*6, r*6+cnt_r), cnt_r in 0..6
    cnt = torch.randint(0, 7, (n,), generator=g)
    start = torch.arange(n) * 6
    se = torch.stack([start, start + cnt], 1).to(torch.int32)
    se[cnt == 0] = -1
    S = n * 6
    d.update(alpha=torch.rand(S, 1, generator=g), rgb=torch.rand(S, 3, generator=g), z=torch.rand(S, 1, generator=g))
    fwd, bwd = _run(se, d["alpha"], d["rgb"], d["z"], d, 1, need_dz=False)
    o = oc.fused_composite_forward(se.numpy(), d["alpha"].numpy(), d["rgb"].numpy(), d["z"].numpy(), dtype=np.float64)
    assert rel_err(fwd[0], o["rgb"]) < TOL and rel_err(fwd[3], o["bgT"]) < TOL
    ob = oc.fused_composite_backward(se.numpy(), d["alpha"].numpy(), d["rgb"].numpy(), d["z"].numpy(), d["g_rgb"].numpy(),
                                     d["g_depth"].numpy(), d["g_acc"].numpy(), d["g_bgT"].numpy())
    own = np.zeros(S, bool)
    for s, e in se.numpy():
        if e > s:
            own[s:e] = True
    assert grad_err(bwd[0][own], ob["d_alpha"][own]) < TOL


def test_autograd_function_matches_dense_torch():
    """CompositeFunc (fused kernels under autograd) vs the dense torch path differentiated by autograd on the CPU."""
    from volsurfs_b200.volume_rendering import composite

    d = dense_layers(3000, 5, seed_offset=21)
    se, a, c, z = pack_dense(d["hit"], d["alpha"], d["rgb"], d["z"])
    a_g = a.cuda().requires_grad_(True)
    c_g = c.cuda().requires_grad_(True)
    rgb, depth, acc, bgT = composite(_rsp(se), a_g, c_g, z.cuda())
    loss = (rgb * d["g_rgb"].cuda()).sum() + (bgT * d["g_bgT"].cuda()).sum() + (depth * d["g_depth"].cuda()).sum()
    loss.backward()
    ad = d["alpha"].clone().requires_grad_(True)
    cd = d["rgb"].clone().requires_grad_(True)
    out = oc.dense_composite_torch(ad, cd, surfs_z=d["z"])
    lo = (out["rgb_fg"] * d["g_rgb"]).sum() + (out["bg_transmittance"] * d["g_bgT"]).sum() + (out["depth"] * d["g_depth"]).sum()
    lo.backward()
    hit = d["hit"].numpy()
    ray, j = np.nonzero(hit[:, ::-1])
    lay = hit.shape[1] - 1 - j
    assert rel_err(rgb.detach().cpu().numpy(), out["rgb_fg"].detach().numpy()) < TOL
    assert rel_err(bgT.detach().cpu().numpy(), out["bg_transmittance"].detach().numpy()) < TOL
    assert grad_err(a_g.grad.cpu().numpy()[:, 0], ad.grad.numpy()[ray, lay, 0]) < TOL
    assert grad_err(c_g.grad.cpu().numpy(), cd.grad.numpy()[ray, lay]) < TOL


def test_full_size_properties():
    """BASELINE-size run (2^22 rays x 5, all hit): size-independent properties instead of an oracle pass —
    acc + bgT == 1, 0 <= bgT <= 1, d_rgb == g_rgb * w, and linearity of the forward in rgb."""
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    n, K = 1 << 22, 5
    d = all_hit_packed(n, K, seed_offset=31)
    rsp = _rsp(d["se"])
    a, c, z = d["alpha"].cuda(), d["rgb"].cuda(), d["z"].cuda()
    rgb, depth, acc, bgT, w, T = VR.composite(rsp, a, c, z, return_weights=True)
    assert torch.allclose(acc + bgT, torch.ones_like(acc), atol=3e-6)
    assert bool((bgT >= 0).all()) and bool((bgT <= 1).all())
    rgb2, *_ = VR.composite(rsp, a, 2 * c, z)
    assert torch.allclose(rgb2, 2 * rgb, rtol=1e-6, atol=1e-7)
    g = {k: d[k].cuda() for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")}
    d_alpha, d_rgb, _ = VR.composite_backward(rsp, a, c, z, g["g_rgb"], g["g_depth"], g["g_acc"], g["g_bgT"])
    ray_of = torch.arange(n, device="cuda").repeat_interleave(K)
    assert torch.allclose(d_rgb, g["g_rgb"][ray_of] * w, rtol=1e-6, atol=1e-7)
    # spot-check 4096 rays against the oracle
    sub = slice(0, 4096 * K)
    o = oc.fused_composite_backward(d["se"][:4096].numpy(), d["alpha"][sub].numpy(), d["rgb"][sub].numpy(), d["z"][sub].numpy(),
                                    d["g_rgb"][:4096].numpy(), d["g_depth"][:4096].numpy(), d["g_acc"][:4096].numpy(),
                                    d["g_bgT"][:4096].numpy())
    assert grad_err(d_alpha[sub].cpu().numpy(), o["d_alpha"]) < TOL
    assert composite_bytes(n, n * K) == n * 344


def test_reference_fp32_error_scale():
    """Calibrates the gradient metric: the reference's own fp32 torch-autograd gradients vs the fp64 truth, and ours vs
    the same truth, under grad_err.  Ours must be within the 1e-5 budget and not worse than 4x the reference's error."""
    from volsurfs_b200.volsurfs import VolumeRendering as VR

    d = dense_layers(4096, 5, seed_offset=1)
    hit = d["hit"].numpy()
    ray, j = np.nonzero(hit[:, ::-1])
    lay = hit.shape[1] - 1 - j

    def dense_grads(dtype):
        a = d["alpha"].to(dtype).clone().requires_grad_(True)
        c = d["rgb"].to(dtype).clone().requires_grad_(True)
        out = oc.dense_composite_torch(a, c)
        ((out["rgb_fg"] * d["g_rgb"].to(dtype)).sum() + (out["bg_transmittance"] * d["g_bgT"].to(dtype)).sum()).backward()
        return a.grad.detach().numpy()[ray, lay, 0], c.grad.detach().numpy()[ray, lay]

    da64, dc64 = dense_grads(torch.float64)
    da32, dc32 = dense_grads(torch.float32)
    se, a, c, z = pack_dense(d["hit"], d["alpha"].detach(), d["rgb"].detach(), d["z"])
    N = se.shape[0]
    g = {"g_rgb": d["g_rgb"], "g_depth": torch.zeros(N, 1), "g_acc": torch.zeros(N, 1), "g_bgT": d["g_bgT"]}
    _, bwd = _run(se, a, c, z, g, 0, need_dz=False)
    ref_err = max(grad_err(da32, da64), grad_err(dc32, dc64))
    our_err = max(grad_err(bwd[0][:, 0], da64), grad_err(bwd[1], dc64))
    print(f"fp32 torch autograd vs fp64: {ref_err:.2e}; CUDA vs fp64: {our_err:.2e}")
    assert our_err < TOL and our_err < 4 * ref_err + 1e-7


def test_scan_kernels_on_unaligned_and_gapped_segments():
    """coarsened scan kernels: ray starts that are not multiples of four, quads shared by neighbouring rays, gaps between rays"""
    g = torch.Generator().manual_seed(17)
    n = 3000
    cnt = torch.randint(0, 40, (n,), generator=g)
    gap = torch.randint(0, 3, (n,), generator=g)
    start = torch.cumsum(cnt + gap, 0) - cnt
    se = torch.stack([start, start + cnt], 1).to(torch.int32)
    se[cnt == 0] = -1
    S = int((start + cnt).max()) + 1   # deliberately not a multiple of four
    d = {"alpha": torch.rand(S, 1, generator=g), "rgb": torch.rand(S, 3, generator=g), "z": torch.rand(S, 1, generator=g),
         "g_rgb": torch.randn(n, 3, generator=g), "g_depth": torch.randn(n, 1, generator=g), "g_acc": torch.randn(n, 1, generator=g),
         "g_bgT": torch.randn(n, 1, generator=g)}
    own = np.zeros(S, bool)
    for s0, e0 in se.numpy():
        if e0 > s0:
            own[s0:e0] = True
    for mode in (2, 3, 5, 6, 7):
        fwd, bwd = _run(se, d["alpha"], d["rgb"], d["z"], d, mode)
        o = oc.fused_composite_forward(se.numpy(), d["alpha"].numpy(), d["rgb"].numpy(), d["z"].numpy(), dtype=np.float64)
        for got, key in zip(fwd[:4], ("rgb", "depth", "acc", "bgT")):
            assert rel_err(got, o[key], floor=1e-6) < TOL, (mode, key)
        assert rel_err(fwd[4][own], o["weights"][own], floor=1e-6) < TOL
        ob = oc.fused_composite_backward(se.numpy(), d["alpha"].numpy(), d["rgb"].numpy(), d["z"].numpy(), d["g_rgb"].numpy(),
                                         d["g_depth"].numpy(), d["g_acc"].numpy(), d["g_bgT"].numpy())
        for got, key in zip(bwd, ("d_alpha", "d_rgb", "d_z")):
            assert grad_err(got[own], ob[key][own]) < TOL, (mode, key)
