"""CPU tests that pin the compositing oracle: against vectors produced by the reference's own source lines
(tests/golden/make_golden.py), the reference's in-code worked example, torch autograd and fp64."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import compositing as oc
from volsurfs_b200.synthetic import dense_layers, nerf_packets


@pytest.mark.parametrize("name", ["dense_composite_k5", "dense_composite_k1", "dense_composite_k9"])
def test_dense_restatement_matches_reference_lines(name):
    g = np.load(GOLDEN / f"{name}.npz")
    alpha, rgb = torch.from_numpy(g["alpha"]), torch.from_numpy(g["rgb"])
    for tag, half in (("fp32", False), ("fp16", True)):
        a = alpha.clone().requires_grad_(True)
        c = rgb.clone().requires_grad_(True)
        out = oc.dense_composite_torch(a, c, rgb_bg=torch.from_numpy(g["rgb_bg"]), half=half)
        # bit-exact: the restatement runs the same torch ops in the same order
        assert np.array_equal(out["rgb_fg"].detach().numpy(), g[f"{tag}_rgb_fg"])
        assert np.array_equal(out["bg_transmittance"].detach().numpy(), g[f"{tag}_bg_transmittance"])
        assert np.array_equal(out["rgb"].detach().numpy(), g[f"{tag}_pred_rgb"])
        assert np.array_equal(out["weights"].detach().numpy(), g[f"{tag}_weights"])
        assert np.array_equal(out["transmittance"].detach().numpy(), g[f"{tag}_transmittance"])
        loss = (out["rgb_fg"] * torch.from_numpy(g["g_rgb"])).sum() + (out["bg_transmittance"] * torch.from_numpy(g["g_bgT"])).sum()
        loss.backward()
        assert np.allclose(a.grad.numpy(), g[f"{tag}_d_alpha"], rtol=0, atol=0)
        assert np.allclose(c.grad.numpy(), g[f"{tag}_d_rgb"], rtol=0, atol=0)


def test_reference_worked_example():
    """VolumeRenderingGPU.cuh:60-62: x=[0.9,0.5,0.1] -> T=[1.0,0.9,0.45], bg=0.45."""
    se = np.array([[0, 3]], np.int32)
    T, bg = oc.packed_cumprod_one_minus_alpha_to_transmittance(se, np.array([[0.9], [0.5], [0.1]], np.float32))
    assert np.allclose(T[:, 0], [1.0, 0.9, 0.45], rtol=1e-7)
    assert np.allclose(bg, [[0.45]], rtol=1e-7)


def test_packed_fused_equals_dense_golden():
    """Packing the golden dense inputs (outer->inner) and running the packed fused oracle reproduces the reference's
    dense outputs and autograd gradients."""
    g = np.load(GOLDEN / "dense_composite_k5.npz")
    hit = g["hit"]
    se, layer, a, c, z = oc.dense_to_packed(hit, g["alpha"], g["rgb"], g["z"])
    f = oc.fused_composite_forward(se, a, c, z)
    assert rel_err(f["rgb"], g["fp32_rgb_fg"]) < 1e-6
    assert rel_err(f["bgT"], g["fp32_bg_transmittance"]) < 1e-6
    N = hit.shape[0]
    b = oc.fused_composite_backward(se, a, c, z, g["g_rgb"], np.zeros((N, 1)), np.zeros((N, 1)), g["g_bgT"])
    ray, j = np.nonzero(hit[:, ::-1])
    lay = hit.shape[1] - 1 - j
    assert np.array_equal(lay, layer)
    assert rel_err(b["d_alpha"][:, 0], g["fp32_d_alpha"][ray, lay, 0], floor=1e-5) < 2e-5
    assert rel_err(b["d_rgb"], g["fp32_d_rgb"][ray, lay], floor=1e-5) < 2e-5


def test_fused_backward_matches_autograd_fp64():
    d = dense_layers(257, 5, seed_offset=7)
    hit = d["hit"].numpy()
    a64 = d["alpha"].double().requires_grad_(True)
    c64 = d["rgb"].double().requires_grad_(True)
    z64 = d["z"].double().requires_grad_(True)
    out = oc.dense_composite_torch(a64, c64, surfs_z=z64)
    loss = ((out["rgb_fg"] * d["g_rgb"].double()).sum() + (out["depth"] * d["g_depth"].double()).sum()
            + (out["acc"] * d["g_acc"].double()).sum() + (out["bg_transmittance"] * d["g_bgT"].double()).sum())
    loss.backward()
    se, layer, a, c, z = oc.dense_to_packed(hit, d["alpha"].numpy(), d["rgb"].numpy(), d["z"].numpy())
    b = oc.fused_composite_backward(se, a, c, z, d["g_rgb"].numpy(), d["g_depth"].numpy(), d["g_acc"].numpy(), d["g_bgT"].numpy())
    ray, j = np.nonzero(hit[:, ::-1])
    assert rel_err(b["d_alpha"][:, 0], a64.grad.numpy()[ray, layer, 0], floor=1e-9) < 1e-9
    assert rel_err(b["d_rgb"], c64.grad.numpy()[ray, layer], floor=1e-9) < 1e-9
    assert rel_err(b["d_z"][:, 0], z64.grad.numpy()[ray, layer, 0], floor=1e-9) < 1e-9


def test_packed_chain_equals_fused_on_nerf_packets():
    """cumprod -> w = alpha*T -> integrate/sum (nerf.py:308-334) agrees with the fused oracle; bgT differs by the
    documented quirk (packed cumprod excludes the last sample)."""
    p = nerf_packets(300, seed_offset=11, max_per_ray=64, mean=12.0)
    se = p["se"].numpy()
    alpha = p["alpha"].numpy()
    T, bg_quirk = oc.packed_cumprod_one_minus_alpha_to_transmittance(se, 1 - alpha)
    w = (alpha * T).astype(np.float32)
    rgb = oc.packed_integrate_with_weights(se, p["rgb"].numpy(), w)
    depth = oc.packed_integrate_with_weights(se, p["z"].numpy(), w)
    acc, acc_ps = oc.packed_sum_over_rays(se, w)
    f = oc.fused_composite_forward(se, alpha, p["rgb"].numpy(), p["z"].numpy())
    assert rel_err(rgb, f["rgb"]) < 1e-6 and rel_err(depth, f["depth"]) < 1e-6 and rel_err(acc, f["acc"]) < 1e-6
    n = (se[:, 1] - se[:, 0])
    last = se[n > 0, 1] - 1
    full = bg_quirk[n > 0, 0] * (1 - alpha[last, 0])
    assert rel_err(full, f["bgT"][n > 0, 0]) < 1e-6
    assert np.all(bg_quirk[n == 0] == 1.0) and np.all(f["bgT"][n == 0] == 1.0)
    # weights + bgT partition unity
    assert np.allclose(f["acc"][:, 0] + f["bgT"][:, 0], 1.0, atol=2e-6)


def test_packed_cumprod_backward_matches_autograd():
    p = nerf_packets(64, seed_offset=12, max_per_ray=40, mean=9.0)
    se = p["se"].numpy()
    x = p["x"].numpy()
    S = x.shape[0]
    rng = np.random.default_rng(0)
    gT = rng.standard_normal((S, 1)).astype(np.float32)
    gbg = rng.standard_normal((se.shape[0], 1)).astype(np.float32)
    T, bg = oc.packed_cumprod_one_minus_alpha_to_transmittance(se, x)
    dx = oc.packed_cumprod_backward_full(se, gT, gbg, x, T, bg)
    # autograd reference, ray by ray, in fp64
    xt = torch.from_numpy(x).double().requires_grad_(True)
    loss = 0
    for r, (s, e) in enumerate(se):
        if e - s == 0:
            continue
        seg = xt[s:e, 0]
        Tr = torch.cat([torch.ones(1, dtype=torch.float64), torch.cumprod(seg[:-1], 0)])
        loss = loss + (Tr * torch.from_numpy(gT[s:e, 0]).double()).sum() + Tr[-1] * float(gbg[r, 0])
    loss.backward()
    assert rel_err(dx, xt.grad.numpy(), floor=1e-4) < 1e-4


def test_sum_cumsum_and_integrate_backward():
    p = nerf_packets(50, seed_offset=13, max_per_ray=20, mean=6.0)
    se = p["se"].numpy()
    S = p["alpha"].shape[0]
    rng = np.random.default_rng(1)
    v = rng.standard_normal((S, 1)).astype(np.float32)
    cs = oc.packed_cumsum_over_rays(se, v, False)
    csr = oc.packed_cumsum_over_rays(se, v, True)
    for s, e in se:
        if e > s:
            assert np.allclose(cs[s:e, 0], np.cumsum(v[s:e, 0]), atol=1e-5)
            assert np.allclose(csr[s:e, 0], np.cumsum(v[s:e, 0][::-1])[::-1], atol=1e-5)
    v3 = rng.standard_normal((S, 3)).astype(np.float32)
    w = rng.random((S, 1)).astype(np.float32)
    g = rng.standard_normal((se.shape[0], 3)).astype(np.float32)
    dv, dw = oc.packed_integrate_with_weights_backward(se, g, v3, w)
    dvb, dwb = oc.packed_integrate_with_weights_backward(se, g, v3, w, ref_bug=True)
    ray_of = np.repeat(np.arange(se.shape[0]), np.maximum(se[:, 1] - se[:, 0], 0))
    assert np.allclose(dv, g[ray_of] * w, atol=1e-6)
    assert np.allclose(dw[:, 0], (g[ray_of] * v3).sum(1), atol=1e-5)
    assert np.array_equal(dv, dvb)
    assert not np.allclose(dw, dwb)  # the reference bug changes dw only
    gs = rng.standard_normal((S, 3)).astype(np.float32)
    d = oc.packed_sum_over_rays_backward(se, g, gs, v3)
    assert np.allclose(d, g[ray_of] + gs, atol=1e-6)
