"""Autograd glue over the packed operators — the counterpart of the reference's
``volsurfs_py/volume_rendering/volume_rendering_funcs.py:91-272`` (four ``autograd.Function``s with the same names
and argument order) and ``volume_rendering_modules.py:62-234`` (``VolumeRendering{,NeRF,NeuS}`` modules), plus the
fused :class:`CompositeFunc` that is the product's fast path.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from .volsurfs import VolumeRendering as _VR


class CumprodOneMinusAlphaToTransmittanceFunc(Function):
    """T_i = prod_{j<i} x_j and bgT (reference funcs.py:91-179). Backward = reverse cumsum of g_T*T divided by
    clamp_min(x,1e-6); one fused launch here instead of mul + cumsum_over_rays + backward kernel."""

    @staticmethod
    def forward(ctx, ray_samples_packed, alpha):
        transmittance, bg_transmittance = _VR.cumprod_one_minus_alpha_to_transmittance(ray_samples_packed, alpha)
        ctx.save_for_backward(alpha, transmittance, bg_transmittance)
        ctx.ray_samples_packed = ray_samples_packed
        return transmittance, bg_transmittance

    @staticmethod
    def backward(ctx, grad_transmittance, grad_bg_transmittance):
        alpha, transmittance, bg_transmittance = ctx.saved_tensors
        rsp = ctx.ray_samples_packed
        grad_alpha = _VR.cumprod_backward_fused(
            grad_transmittance.contiguous(), grad_bg_transmittance.contiguous(), rsp, alpha, transmittance, bg_transmittance
        )
        ctx.ray_samples_packed = None
        return None, grad_alpha


class IntegrateWithWeights1DFunc(Function):
    @staticmethod
    def forward(ctx, ray_samples_packed, values_samples, weights_samples):
        out = _VR.integrate_with_weights_1d(ray_samples_packed, values_samples, weights_samples)
        ctx.save_for_backward(values_samples, weights_samples, out)
        ctx.ray_samples_packed = ray_samples_packed
        return out

    @staticmethod
    def backward(ctx, grad_out):
        values_samples, weights_samples, out = ctx.saved_tensors
        dv, dw = _VR.integrate_with_weights_1d_backward(grad_out.contiguous(), ctx.ray_samples_packed, values_samples, weights_samples, out)
        ctx.ray_samples_packed = None
        return None, dv, dw


class IntegrateWithWeights3DFunc(Function):
    @staticmethod
    def forward(ctx, ray_samples_packed, values_samples, weights_samples):
        out = _VR.integrate_with_weights_3d(ray_samples_packed, values_samples, weights_samples)
        ctx.save_for_backward(values_samples, weights_samples, out)
        ctx.ray_samples_packed = ray_samples_packed
        return out

    @staticmethod
    def backward(ctx, grad_out):
        values_samples, weights_samples, out = ctx.saved_tensors
        dv, dw = _VR.integrate_with_weights_3d_backward(grad_out.contiguous(), ctx.ray_samples_packed, values_samples, weights_samples, out)
        ctx.ray_samples_packed = None
        return None, dv, dw


class SumOverRayFunc(Function):
    @staticmethod
    def forward(ctx, ray_samples_packed, sample_values):
        per_ray, per_sample = _VR.sum_over_rays(ray_samples_packed, sample_values)
        ctx.save_for_backward(sample_values)
        ctx.ray_samples_packed = ray_samples_packed
        return per_ray, per_sample

    @staticmethod
    def backward(ctx, grad_per_ray, grad_per_sample):
        (sample_values,) = ctx.saved_tensors
        dv = _VR.sum_over_rays_backward(grad_per_ray.contiguous(), grad_per_sample.contiguous(), ctx.ray_samples_packed, sample_values)
        ctx.ray_samples_packed = None
        return None, dv


class CompositeFunc(Function):
    """Fused compositing: (rgb [N,3], depth [N,1], acc [N,1], bgT [N,1]) = f(alpha [S,1], rgb [S,3], z [S,1]).

    Equals the dense K-layer torch path of the volsurfs method (volsurfs.py:601-640,708) on packed layer hits and
    the chain of nerf.py:308-334 on NeRF packets (with bgT as the full product).  Nothing but the inputs is saved."""

    @staticmethod
    def forward(ctx, ray_samples_packed, alpha, rgb, z):
        out = _VR.composite(ray_samples_packed, alpha, rgb, z)
        ctx.save_for_backward(alpha, rgb, z)
        ctx.ray_samples_packed = ray_samples_packed
        return out

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, g_bgT):
        alpha, rgb, z = ctx.saved_tensors
        rsp = ctx.ray_samples_packed
        n = rsp.get_nr_rays()

        def z0(g, cols):
            return torch.zeros((n, cols), dtype=torch.float32, device=alpha.device) if g is None else g.contiguous()

        need_dz = ctx.needs_input_grad[3]
        d_alpha, d_rgb, d_z = _VR.composite_backward(
            rsp, alpha, rgb, z, z0(g_rgb, 3), z0(g_depth, 1), z0(g_acc, 1), z0(g_bgT, 1), need_dz=need_dz
        )
        ctx.ray_samples_packed = None
        return None, d_alpha, d_rgb, d_z


def composite(ray_samples_packed, alpha, rgb, z=None):
    """Differentiable fused compositing; ``z`` defaults to the packet's ``samples_z``."""
    if z is None:
        z = ray_samples_packed.samples_z
    return CompositeFunc.apply(ray_samples_packed, alpha, rgb, z)


# ---- nn.Module wrappers (reference volume_rendering_modules.py) -------------------------------------------------------
class CumprodOneMinusAlphaToTransmittanceModule(torch.nn.Module):
    def forward(self, ray_samples_packed, alpha):
        return CumprodOneMinusAlphaToTransmittanceFunc.apply(ray_samples_packed, alpha)


class IntegrateWithWeights1DModule(torch.nn.Module):
    def forward(self, ray_samples_packed, values_samples, weights_samples):
        return IntegrateWithWeights1DFunc.apply(ray_samples_packed, values_samples, weights_samples)


class IntegrateWithWeights3DModule(torch.nn.Module):
    def forward(self, ray_samples_packed, values_samples, weights_samples):
        return IntegrateWithWeights3DFunc.apply(ray_samples_packed, values_samples, weights_samples)


class SumOverRayModule(torch.nn.Module):
    def forward(self, ray_samples_packed, sample_values):
        return SumOverRayFunc.apply(ray_samples_packed, sample_values)


class VolumeRendering(torch.nn.Module):
    """reference volume_rendering_modules.py:62-85"""

    def __init__(self):
        super().__init__()
        self.cumprod_one_minus_alpha_to_transmittance_module = CumprodOneMinusAlphaToTransmittanceModule()
        self.integrator_1d_module = IntegrateWithWeights1DModule()
        self.integrator_3d_module = IntegrateWithWeights3DModule()
        self.sum_ray_module = SumOverRayModule()

    def integrate_1d(self, ray_samples_packed, samples_vals, weights):
        assert samples_vals.shape[1] == 1, "samples_vals should be 1d"
        return self.integrator_1d_module(ray_samples_packed, samples_vals, weights)

    def integrate_3d(self, ray_samples_packed, samples_vals, weights):
        assert samples_vals.shape[1] == 3, "samples_vals should be 3d"
        return self.integrator_3d_module(ray_samples_packed, samples_vals, weights)

    def composite(self, ray_samples_packed, alpha, rgb, z=None):
        return composite(ray_samples_packed, alpha, rgb, z)


class VolumeRenderingNeRF(VolumeRendering):
    """reference volume_rendering_modules.py:88-106: alpha = clamp(1-exp(-sigma*dt)), x = 1-alpha+1e-6"""

    def compute_weights(self, ray_samples_packed, samples_densities):
        dt = ray_samples_packed.samples_dt
        alpha = torch.clamp(1.0 - torch.exp(-samples_densities * dt), min=0.0, max=1.0)
        transmittance, bg_transmittance = self.cumprod_one_minus_alpha_to_transmittance_module(
            ray_samples_packed, 1 - alpha + 1e-6
        )
        return alpha * transmittance, bg_transmittance


class VolumeRenderingNeuS(VolumeRendering):
    """reference volume_rendering_modules.py:109-234"""

    def compute_alphas_from_logistic_beta(self, ray_samples_packed, sdf, gradients, cos_anneal_ratio, logistic_beta, debug_ray_idx=None):
        """NeuS discrete opacity (modules.py:115-216): section-point SDFs from the annealed ray/normal cosine, logistic CDF, ratio"""
        dists = ray_samples_packed.samples_dt
        true_cos = (ray_samples_packed.samples_dirs * gradients).sum(-1, keepdim=True)
        relu = torch.nn.functional.relu
        iter_cos = -(relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + relu(-true_cos) * cos_anneal_ratio)
        # (iter_cos * dists) * 0.5, in that order: the rounding sequence of the reference expression
        prev_cdf = torch.sigmoid((sdf - iter_cos * dists * 0.5) * logistic_beta)
        next_cdf = torch.sigmoid((sdf + iter_cos * dists * 0.5) * logistic_beta)
        return ((prev_cdf - next_cdf + 1e-6) / (prev_cdf + 1e-6)).clip(0.0, 1.0)

    def compute_transmittance_from_alphas(self, ray_samples_packed, alpha):
        transmittance, _ = self.cumprod_one_minus_alpha_to_transmittance_module(ray_samples_packed, (1 - alpha) + 1e-6)
        return transmittance

    def compute_weights_from_transmittance_and_alphas(self, ray_samples_packed, transmittance, alpha):
        return (alpha * transmittance).view(-1, 1)
