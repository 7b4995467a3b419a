"""Golden vectors for the autograd glue, produced by RUNNING THE REFERENCE'S OWN, UNMODIFIED GLUE FILES
(volsurfs_py/volume_rendering/volume_rendering_funcs.py:91-272 and volume_rendering_modules.py:62-234, imported from /root/reference)
on CPU tensors.  Those files import `from volsurfs import VolumeRendering` — the pybind module, GPU-only in the reference; here a CPU
stand-in module named `volsurfs` serves the nine static operators they call from the packed-operator oracle (oracle/compositing.py,
itself pinned to the reference's kernels by tests/test_gpu_reference_kernels.py).  Run where /root/reference is mounted; the .npz is
committed, the reference sources are not.

    python tests/golden/make_golden_glue.py

glue_nerf_neus.npz: a config-C3-shaped packet (1024 rays, 35 % empty, lognormal counts <= 128), inputs, and for
  * the NeRF chain of methods/nerf.py:308-334 through VolumeRenderingNeRF.compute_weights + SumOverRayFunc + integrate_3d + integrate_1d
  * the NeuS chain of methods/surf.py:383-428 through VolumeRenderingNeuS.compute_alphas_from_logistic_beta / compute_transmittance_from_alphas
    / compute_weights_from_transmittance_and_alphas + integrate_3d
the outputs and the autograd gradients the reference's Functions return.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
ROOT = OUT.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import compositing as oc  # noqa: E402
from volsurfs_b200.synthetic import nerf_packets  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class VolumeRendering:
    """CPU stand-in for the pybind class (src/PyBridge.cxx:113-129): same static methods, argument order and shapes"""

    @staticmethod
    def cumprod_one_minus_alpha_to_transmittance(rsp, x):
        T, bg = oc.packed_cumprod_one_minus_alpha_to_transmittance(_np(rsp.ray_start_end_idx), _np(x))
        return _t(T), _t(bg)

    @staticmethod
    def cumsum_over_rays(rsp, v, inverse):
        return _t(oc.packed_cumsum_over_rays(_np(rsp.ray_start_end_idx), _np(v), inverse))

    @staticmethod
    def cumprod_one_minus_alpha_to_transmittance_backward(gT, gbg, rsp, x, T, bg, cumsumLV):
        return _t(oc.packed_cumprod_backward(_np(rsp.ray_start_end_idx), _np(gT), _np(gbg), _np(x), _np(T), _np(bg), _np(cumsumLV)))

    @staticmethod
    def integrate_with_weights_1d(rsp, v, w):
        return _t(oc.packed_integrate_with_weights(_np(rsp.ray_start_end_idx), _np(v), _np(w)))

    integrate_with_weights_3d = integrate_with_weights_1d

    @staticmethod
    def integrate_with_weights_1d_backward(g, rsp, v, w, result):
        dv, dw = oc.packed_integrate_with_weights_backward(_np(rsp.ray_start_end_idx), _np(g), _np(v), _np(w), ref_bug=False)
        return _t(dv), _t(dw)

    integrate_with_weights_3d_backward = integrate_with_weights_1d_backward  # mathematically correct dw (the product's default mode)

    @staticmethod
    def sum_over_rays(rsp, v):
        per_ray, per_sample = oc.packed_sum_over_rays(_np(rsp.ray_start_end_idx), _np(v))
        return _t(per_ray), _t(per_sample)

    @staticmethod
    def sum_over_rays_backward(g_ray, g_sample, rsp, v):
        return _t(oc.packed_sum_over_rays_backward(_np(rsp.ray_start_end_idx), _np(g_ray), _np(g_sample), _np(v)))


def import_reference_glue():
    """the reference's files, unmodified, from where they lie"""
    mod = types.ModuleType("volsurfs")
    mod.VolumeRendering = VolumeRendering
    sys.modules["volsurfs"] = mod
    sys.path.insert(0, "/root/reference")
    from volsurfs_py.volume_rendering import volume_rendering_funcs as funcs  # noqa: E402
    from volsurfs_py.volume_rendering import volume_rendering_modules as modules  # noqa: E402

    assert funcs.__file__.startswith("/root/reference/")
    return funcs, modules


def inputs(n_rays=1024, seed_offset=77):
    p = nerf_packets(n_rays, seed_offset=seed_offset, max_per_ray=128, mean=24.0)
    S = p["alpha"].shape[0]
    g = torch.Generator().manual_seed(4242)
    return {
        "se": p["se"], "dt": p["dt"], "z": p["z"],
        "density": (torch.rand(S, 1, generator=g) * 40.0) * (torch.rand(S, 1, generator=g) > 0.5),
        "rgb": torch.rand(S, 3, generator=g),
        "dirs": torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=1),
        "sdf": torch.randn(S, 1, generator=g) * 0.05,
        "gradients": torch.nn.functional.normalize(torch.randn(S, 3, generator=g), dim=1),
        "g_rgb": torch.randn(n_rays, 3, generator=g), "g_depth": torch.randn(n_rays, 1, generator=g),
        "g_wsum": torch.randn(n_rays, 1, generator=g), "g_bgT": torch.randn(n_rays, 1, generator=g),
    }


def run_chains(funcs, modules, x, rsp):
    """the two call sequences; `rsp` only needs ray_start_end_idx / samples_dt / samples_dirs.  Returns a flat dict of tensors."""
    out = {}
    # ---- NeRF: methods/nerf.py:308-334 ------------------------------------------------------------------------------
    vr = modules.VolumeRenderingNeRF()
    dens = x["density"].clone().requires_grad_(True)
    rgb = x["rgb"].clone().requires_grad_(True)
    weights, bgT = vr.compute_weights(rsp, dens)
    wsum, _ = funcs.SumOverRayFunc.apply(rsp, weights)
    pred_rgb = vr.integrate_3d(rsp, rgb, weights)
    pred_depth = vr.integrate_1d(rsp, x["z"], weights)
    loss = (pred_rgb * x["g_rgb"]).sum() + (pred_depth * x["g_depth"]).sum() + (wsum * x["g_wsum"]).sum() + (bgT * x["g_bgT"]).sum()
    loss.backward()
    out.update(nerf_weights=weights, nerf_bgT=bgT, nerf_wsum=wsum, nerf_rgb=pred_rgb, nerf_depth=pred_depth, nerf_d_density=dens.grad,
               nerf_d_rgb=rgb.grad)
    # ---- NeuS: methods/surf.py:383-428 ------------------------------------------------------------------------------
    vn = modules.VolumeRenderingNeuS()
    sdf = x["sdf"].clone().requires_grad_(True)
    rgb2 = x["rgb"].clone().requires_grad_(True)
    alpha = vn.compute_alphas_from_logistic_beta(rsp, sdf, x["gradients"], 0.7, 64.0)
    T = vn.compute_transmittance_from_alphas(rsp, alpha)
    w = vn.compute_weights_from_transmittance_and_alphas(rsp, T, alpha)
    col = vn.integrate_3d(rsp, rgb2, w)
    (col * x["g_rgb"]).sum().backward()
    out.update(neus_alpha=alpha, neus_T=T, neus_weights=w, neus_rgb=col, neus_d_sdf=sdf.grad, neus_d_rgb=rgb2.grad)
    return out


def main():
    funcs, modules = import_reference_glue()
    x = inputs()
    rsp = types.SimpleNamespace(ray_start_end_idx=x["se"], samples_dt=x["dt"], samples_dirs=x["dirs"])
    out = run_chains(funcs, modules, x, rsp)
    np.savez_compressed(OUT / "glue_nerf_neus.npz", **{f"in_{k}": _np(v) for k, v in x.items()}, **{k: _np(v) for k, v in out.items()})
    print({k: (tuple(v.shape), float(v.abs().mean())) for k, v in out.items()})


if __name__ == "__main__":
    main()
