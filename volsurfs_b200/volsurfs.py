"""Python twin of the reference's pybind11 module ``volsurfs`` (src/PyBridge.cxx:19) for the hot path.

Same class names, method names, argument order and tensor shapes as ``PyBridge.cxx:70-129`` so that
``volsurfs_py/volume_rendering/*`` and the renderers can ``from volsurfs import VolumeRendering,
RaySamplesPacked`` unchanged (see :func:`volsurfs_b200.install_as_volsurfs`).  Every operator goes through the
C ABI of ``libvolsurfs_b200.so`` on the caller's current CUDA stream; there is no CPU or PyTorch fallback.

Deliberate differences from the reference:
  * tensors live on the *current* CUDA device instead of a hard-coded ``cuda:0``
    (src/VolumeRendering.cu:45, src/RaySamplesPacked.cu:16-42);
  * kernels are enqueued on torch's current stream and nothing calls ``cudaDeviceSynchronize``
    (src/VolumeRendering.cu:54,65);
  * argument errors raise ``RuntimeError``/``ValueError`` instead of aborting the process through ``CHECK``;
  * ``integrate_with_weights_3d_backward`` computes the mathematically correct ``dw`` unless
    ``VolumeRendering.reference_bugs = True`` (the reference reads channel [1] twice,
    kernels/volsurfs/VolumeRenderingGPU.cuh:1021).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.VolsurfsB200Error("volsurfs_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, name: str, cols=None) -> torch.Tensor:
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 2:
        raise RuntimeError(f"{name} must be 2-dimensional, got shape {tuple(t.shape)}")
    if cols is not None and t.shape[1] != cols:
        raise RuntimeError(f"{name} must have {cols} columns, got shape {tuple(t.shape)}")
    return t.contiguous()


class RaySamplesPacked:
    """Struct-of-arrays packet of per-ray sample segments (include/volsurfs/RaySamplesPacked.cuh:7-81)."""

    def __init__(self, nr_rays: int, max_nr_samples: int, first_sample_idx: int = 0, values_dim: int = 1):
        dev = _device()
        f = dict(dtype=torch.float32, device=dev)
        # per sample (src/RaySamplesPacked.cu:15-24)
        self.samples_idx = torch.arange(
            first_sample_idx, max_nr_samples + first_sample_idx, dtype=torch.int32, device=dev
        ).unsqueeze(1)
        self.samples_3d = torch.full((max_nr_samples, 3), -1.0, **f)
        self.samples_dirs = torch.full((max_nr_samples, 3), -1.0, **f)
        self.samples_z = torch.full((max_nr_samples, 1), -1.0, **f)
        self.samples_dt = torch.full((max_nr_samples, 1), -1.0, **f)
        self.samples_values = torch.full((max_nr_samples, values_dim), -1.0, **f)
        # per ray (:26-38)
        self.ray_start_end_idx = torch.full((nr_rays, 2), -1, dtype=torch.int32, device=dev)
        self.ray_o = torch.full((nr_rays, 3), -1.0, **f)
        self.ray_d = torch.full((nr_rays, 3), -1.0, **f)
        self.ray_enter = torch.full((nr_rays, 1), -1.0, **f)
        self.ray_exit = torch.full((nr_rays, 1), -1.0, **f)
        self.ray_max_dt = torch.full((nr_rays, 1), -1.0, **f)
        # flags (:45-47); C++-only in the reference, plain attributes here
        self.has_samples_values = False
        self.has_dt = False
        self.is_compacted = True
        # extension: per-sample layer / triangle / barycentric (u,v) of K-layer hits (None for other producers)
        self.samples_layer = None
        self.samples_triangle = None
        self.samples_uv = None

    @classmethod
    def _from_tensors(cls, **tensors) -> "RaySamplesPacked":
        self = cls.__new__(cls)
        self.has_samples_values = False
        self.has_dt = False
        self.is_compacted = True
        self.samples_layer = self.samples_triangle = self.samples_uv = None
        for k, v in tensors.items():
            setattr(self, k, v)
        return self

    # -- sizes ---------------------------------------------------------------------------------------------------
    def get_nr_rays(self) -> int:
        return int(self.ray_start_end_idx.shape[0])

    def get_max_nr_samples(self) -> int:
        return int(self.samples_idx.shape[0])

    def get_values_dim(self) -> int:
        return int(self.samples_values.shape[1])

    def get_nr_samples_per_ray(self) -> torch.Tensor:
        se = self.ray_start_end_idx
        return se[:, 1] - se[:, 0]

    def get_total_nr_samples(self) -> int:
        """Sum of the per-ray counts (src/RaySamplesPacked.cu:170-175). One device->host read."""
        n = self.get_nr_rays()
        if n == 0:
            return 0
        total = torch.empty(1, dtype=torch.int64, device=self.ray_start_end_idx.device)
        check(_lib.lib().vs_count_total(ptr(self.ray_start_end_idx), n, ptr(total), _stream()), "vs_count_total")
        return int(total.item())

    def is_empty(self) -> bool:
        return self.get_nr_rays() == 0 or self.get_total_nr_samples() == 0

    # -- per-ray accessors (src/RaySamplesPacked.cu:55-150) ---------------------------------------------------------
    def _check_ray(self, ray_idx: int):
        if ray_idx < 0 or ray_idx >= self.get_nr_rays():
            raise ValueError("ray_idx must be in the range [0, nr_rays)")

    def get_ray_max_dt(self, ray_idx: int) -> float:
        self._check_ray(ray_idx)
        return float(self.ray_max_dt[ray_idx, 0].item())

    def _ray_slice(self, t: torch.Tensor, ray_idx: int) -> torch.Tensor:
        s, e = self.ray_start_end_idx[ray_idx].tolist()
        return t[s:e]

    def get_ray_samples_idx(self, ray_idx: int):
        return self._ray_slice(self.samples_idx, ray_idx)

    def get_ray_samples_3d(self, ray_idx: int):
        return self._ray_slice(self.samples_3d, ray_idx)

    def get_ray_samples_dirs(self, ray_idx: int):
        return self._ray_slice(self.samples_dirs, ray_idx)

    def get_ray_samples_z(self, ray_idx: int):
        return self._ray_slice(self.samples_z, ray_idx)

    def get_ray_samples_dt(self, ray_idx: int):
        return self._ray_slice(self.samples_dt, ray_idx)

    def get_samples_values(self):
        return self.samples_values

    def get_ray_samples_values(self, ray_idx: int):
        if not self.has_samples_values:
            raise RuntimeError("RaySamplesPacked does not have samples values")
        return self._ray_slice(self.samples_values, ray_idx)

    def get_ray_start_end_idx(self, ray_idx: int):
        return self.ray_start_end_idx[ray_idx, :]

    def get_ray_o(self, ray_idx: int):
        return self.ray_o[ray_idx, :]

    def get_ray_d(self, ray_idx: int):
        return self.ray_d[ray_idx, :]

    def get_ray_enter(self, ray_idx: int):
        return self.ray_enter[ray_idx, :]

    def get_ray_exit(self, ray_idx: int):
        return self.ray_exit[ray_idx, :]

    # -- values ----------------------------------------------------------------------------------------------------
    def are_samples_values_set(self) -> bool:
        return self.has_samples_values

    def set_samples_values(self, samples_values: torch.Tensor) -> None:
        # src/RaySamplesPacked.cu:328-340
        if not self.is_compacted:
            raise RuntimeError("RaySamplesPacked must be compacted before calling set_samples_values")
        if self.has_samples_values:
            raise RuntimeError("Trying to set samples_values when it is already set, remove them first")
        if samples_values.dim() != 2:
            raise RuntimeError("samples_values must be 2-dimensional")
        if self.samples_3d.shape[0] != samples_values.shape[0]:
            raise RuntimeError("samples_3d and samples_values do not have matching 0 dimension")
        self.samples_values = samples_values.clone()
        self.has_samples_values = True

    def remove_samples_values(self) -> None:
        if not self.has_samples_values:
            raise RuntimeError("trying to remove samples_values when it is not set")
        self.samples_values.fill_(-1.0)
        self.has_samples_values = False

    def copy(self) -> "RaySamplesPacked":
        # src/RaySamplesPacked.cu:300-326
        out = RaySamplesPacked._from_tensors(
            samples_idx=self.samples_idx.clone(),
            samples_3d=self.samples_3d.clone(),
            samples_dirs=self.samples_dirs.clone(),
            samples_z=self.samples_z.clone(),
            samples_dt=self.samples_dt.clone(),
            samples_values=self.samples_values.clone() if self.has_samples_values else torch.full_like(self.samples_values, -1.0),
            ray_start_end_idx=self.ray_start_end_idx.clone(),
            ray_o=self.ray_o.clone(),
            ray_d=self.ray_d.clone(),
            ray_enter=self.ray_enter.clone(),
            ray_exit=self.ray_exit.clone(),
            ray_max_dt=self.ray_max_dt.clone(),
        )
        out.has_samples_values = self.has_samples_values
        out.has_dt = self.has_dt
        out.is_compacted = self.is_compacted
        for k in ("samples_layer", "samples_triangle", "samples_uv"):
            v = getattr(self, k)
            setattr(out, k, None if v is None else v.clone())
        return out

    # -- packing ---------------------------------------------------------------------------------------------------
    def compact_to_valid_samples(self) -> "RaySamplesPacked":
        """Gather every ray's segment to dense prefix-sum offsets (src/RaySamplesPacked.cu:188-273).

        One device->host read (the total, needed to size the outputs exactly like the reference); the reference
        takes three (``get_total_nr_samples`` twice + ``is_empty``)."""
        if self.is_compacted:
            raise RuntimeError("RaySamplesPacked must not be compacted before calling compact_to_valid_samples")
        L = _lib.lib()
        n_rays = self.get_nr_rays()
        dev = self.ray_start_end_idx.device
        st = _stream()
        se_in = self.ray_start_end_idx.contiguous()
        scratch = torch.empty(max(int(L.vs_pack_scratch_bytes(n_rays)), 8), dtype=torch.uint8, device=dev)
        total_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.vs_compact_offsets(ptr(se_in), n_rays, ptr(total_dev), ptr(scratch), st), "vs_compact_offsets")
        total = int(total_dev.item())
        if total > 2**31 - 1:
            raise RuntimeError("more than 2^31-1 samples cannot be indexed by int32 ray_start_end_idx")
        vd = self.get_values_dim()
        f = dict(dtype=torch.float32, device=dev)
        out = RaySamplesPacked._from_tensors(
            samples_idx=torch.empty((total, 1), dtype=torch.int32, device=dev),
            samples_3d=torch.empty((total, 3), **f),
            samples_dirs=torch.empty((total, 3), **f),
            samples_z=torch.empty((total, 1), **f),
            samples_dt=torch.empty((total, 1), **f),
            samples_values=torch.empty((total, vd), **f),
            ray_start_end_idx=torch.empty((n_rays, 2), dtype=torch.int32, device=dev),
            ray_o=self.ray_o.clone(),
            ray_d=self.ray_d.clone(),
            ray_enter=self.ray_enter.clone(),
            ray_exit=self.ray_exit.clone(),
            ray_max_dt=self.ray_max_dt.clone(),
        )
        out.has_samples_values = self.has_samples_values
        out.has_dt = self.has_dt
        out.is_compacted = True
        if n_rays == 0:
            return out
        check(
            L.vs_compact_gather(
                ptr(se_in), ptr(scratch),
                ptr(self.samples_idx.contiguous()), ptr(self.samples_3d.contiguous()), ptr(self.samples_dirs.contiguous()),
                ptr(self.samples_z.contiguous()), ptr(self.samples_dt.contiguous()), ptr(self.samples_values.contiguous()), vd,
                ptr(out.ray_start_end_idx), ptr(out.samples_idx), ptr(out.samples_3d), ptr(out.samples_dirs),
                ptr(out.samples_z), ptr(out.samples_dt), ptr(out.samples_values), n_rays, total, st,
            ),
            "vs_compact_gather",
        )
        return out

    def update_dt(self, is_background: bool) -> None:
        """dt_i = clamp(z_{i+1}-z_i, 0, max_dt); last = clamp(t_exit-z, 0, max_dt) or 1e10 for the background
        (kernels/volsurfs/RaySamplesPackedGPU.cuh:14-88)."""
        if not self.is_compacted:
            raise RuntimeError("RaySamplesPacked must be compacted before calling update_dt")
        n_rays = self.get_nr_rays()
        n_samples = self.get_max_nr_samples()
        if n_rays == 0 or n_samples == 0:
            raise RuntimeError("RaySamplesPacked must not be empty before calling update_dt")
        self.samples_dt = self.samples_dt.contiguous()
        check(
            _lib.lib().vs_update_dt(
                ptr(self.ray_start_end_idx), ptr(self.samples_z.contiguous()), ptr(self.ray_exit.contiguous()),
                ptr(self.ray_max_dt.contiguous()), ptr(self.samples_dt), int(bool(is_background)), n_rays, n_samples, _stream(),
            ),
            "vs_update_dt",
        )
        self.has_dt = True


class VolumeRendering:
    """Static packed-sample operators (include/volsurfs/VolumeRendering.cuh:19-97, bound at PyBridge.cxx:113-129)."""

    #: True reproduces the reference's two indexing bugs: integrate_with_weights_3d_backward reads channel [1] twice
    #: (VolumeRenderingGPU.cuh:1021) and median_depth_over_rays' fallback reads samples_z[nr_samples-1] without idx_start (:407)
    reference_bugs = False

    @staticmethod
    def _prep(rsp: RaySamplesPacked, what: str):
        if not rsp.is_compacted:
            raise RuntimeError(f"RaySamplesPacked must be compacted before calling {what}")
        se = rsp.ray_start_end_idx
        if se.dtype != torch.int32 or se.dim() != 2 or se.shape[1] != 2 or not se.is_cuda:
            raise RuntimeError("ray_start_end_idx must be an int32 CUDA tensor of shape [nr_rays, 2]")
        return se.contiguous(), int(se.shape[0])

    @staticmethod
    def _rows(rsp: RaySamplesPacked, *tensors):
        """per-sample tensors must cover the packet's sample slots (the kernels index start+i without a bound; the reference guards
        idx >= max_nr_samples in its loops)"""
        need = rsp.get_max_nr_samples()
        for t in tensors:
            if t is not None and t.shape[0] < need:
                raise RuntimeError(f"per-sample tensor has {t.shape[0]} rows, the packet addresses {need} samples")

    # ---- forward ops ---------------------------------------------------------------------------------------------
    @staticmethod
    def cumprod_one_minus_alpha_to_transmittance(ray_samples_packed, alpha_samples):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "cumprod")
        x = _f32c(alpha_samples, "alpha_samples", 1)
        VolumeRendering._rows(ray_samples_packed, x)
        S = x.shape[0]
        T = torch.zeros((S, 1), dtype=torch.float32, device=x.device)
        bg = torch.empty((n_rays, 1), dtype=torch.float32, device=x.device)
        check(_lib.lib().vs_cumprod_fwd(ptr(se), ptr(x), ptr(T), ptr(bg), n_rays, S, _stream()), "vs_cumprod_fwd")
        return T, bg

    @staticmethod
    def _integrate(ray_samples_packed, values, weights, dim):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, f"integrate_with_weights_{dim}d")
        v = _f32c(values, "values", dim)
        w = _f32c(weights, "weights", 1)
        if v.shape[0] != w.shape[0]:
            raise RuntimeError("values and weights must have the same number of samples")
        VolumeRendering._rows(ray_samples_packed, v)
        out = torch.empty((n_rays, dim), dtype=torch.float32, device=v.device)
        check(_lib.lib().vs_integrate_fwd(ptr(se), ptr(v), ptr(w), ptr(out), dim, n_rays, v.shape[0], _stream()), "vs_integrate_fwd")
        return out

    @staticmethod
    def integrate_with_weights_1d(ray_samples_packed, values, weights):
        return VolumeRendering._integrate(ray_samples_packed, values, weights, 1)

    @staticmethod
    def integrate_with_weights_3d(ray_samples_packed, values, weights):
        return VolumeRendering._integrate(ray_samples_packed, values, weights, 3)

    @staticmethod
    def sum_over_rays(ray_samples_packed, samples_values):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "sum_over_rays")
        v = _f32c(samples_values, "samples_values")
        VolumeRendering._rows(ray_samples_packed, v)
        d = v.shape[1]
        if not (d <= 3 or d == 32) or d == 0:
            raise RuntimeError(f"samples_values should have 1, 2, 3 or 32 values per sample, got {tuple(v.shape)}")
        per_ray = torch.empty((n_rays, d), dtype=torch.float32, device=v.device)
        per_sample = torch.zeros_like(v)
        check(_lib.lib().vs_sum_fwd(ptr(se), ptr(v), ptr(per_ray), ptr(per_sample), d, n_rays, v.shape[0], _stream()), "vs_sum_fwd")
        return per_ray, per_sample

    @staticmethod
    def cumsum_over_rays(ray_samples_packed, samples_values, inverse):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "cumsum_over_rays")
        v = _f32c(samples_values, "samples_values", 1)
        VolumeRendering._rows(ray_samples_packed, v)
        out = torch.zeros_like(v)
        check(_lib.lib().vs_cumsum(ptr(se), ptr(v), ptr(out), int(bool(inverse)), n_rays, v.shape[0], _stream()), "vs_cumsum")
        return out

    # ---- "next" ops of the container: NeuS alpha, median depth, cdf -------------------------------------------
    @staticmethod
    def sdf2alpha(ray_samples_packed, samples_sdf, logistic_beta):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "sdf2alpha")
        if not ray_samples_packed.has_dt:
            raise RuntimeError("ray_samples_packed should have dt")  # CHECK at VolumeRendering.cu:188
        sdf = _f32c(samples_sdf, "samples_sdf", 1)
        beta = _f32c(logistic_beta, "logistic_beta", 1)
        dt = _f32c(ray_samples_packed.samples_dt, "samples_dt", 1)
        VolumeRendering._rows(ray_samples_packed, sdf, beta, dt)
        alpha = torch.zeros_like(sdf)
        check(_lib.lib().vs_sdf2alpha(ptr(se), ptr(dt), ptr(sdf), ptr(beta), ptr(alpha), n_rays, sdf.shape[0], _stream()), "vs_sdf2alpha")
        return alpha

    @staticmethod
    def median_depth_over_rays(ray_samples_packed, samples_weights, threshold):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "median_depth")
        w = _f32c(samples_weights, "samples_weights", 1)
        z = _f32c(ray_samples_packed.samples_z, "samples_z", 1)
        VolumeRendering._rows(ray_samples_packed, w, z)
        out = torch.zeros((n_rays, 1), dtype=torch.float32, device=w.device)
        check(
            _lib.lib().vs_median_depth(ptr(se), ptr(z), ptr(w), float(threshold), ptr(out), n_rays, w.shape[0],
                                       int(VolumeRendering.reference_bugs), _stream()),
            "vs_median_depth",
        )
        return out

    @staticmethod
    def compute_cdf(ray_samples_packed, samples_weights):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "compute_cdf")
        w = _f32c(samples_weights, "samples_weights", 1)
        VolumeRendering._rows(ray_samples_packed, w)
        cdf = torch.zeros_like(w)
        check(_lib.lib().vs_compute_cdf(ptr(se), ptr(w), ptr(cdf), n_rays, w.shape[0], _stream()), "vs_compute_cdf")
        return cdf

    # ---- importance sampling chain ----------------------------------------------------------------------
    #: host copy of the reference's static ``pcg32 m_rng`` (kernels/volsurfs/pcg32.h:32-34 defaults), passed by value to the
    #: jittered launch and advanced by 2^32 afterwards (src/VolumeRendering.cu:520-523)
    _rng_state = 0x853C49E6748FEA9B
    _rng_inc = 0xDA3E39CB94B95BDB

    @staticmethod
    def _rng_advance(delta: int = 1 << 32) -> None:
        m64 = (1 << 64) - 1
        cur_mult, cur_plus, acc_mult, acc_plus = 0x5851F42D4C957F2D, VolumeRendering._rng_inc, 1, 0
        delta &= m64
        while delta > 0:
            if delta & 1:
                acc_mult = (acc_mult * cur_mult) & m64
                acc_plus = (acc_plus * cur_mult + cur_plus) & m64
            cur_plus = ((cur_mult + 1) * cur_plus) & m64
            cur_mult = (cur_mult * cur_mult) & m64
            delta >>= 1
        VolumeRendering._rng_state = (acc_mult * VolumeRendering._rng_state + acc_plus) & m64

    @staticmethod
    def importance_sample(ray_samples_packed, samples_cdf, nr_importance_samples, jitter_samples):
        """src/VolumeRendering.cu:466-548: ``nr_importance_samples`` new samples per ray placed by inverting the per-ray cdf,
        returned as a compacted packet (samples_idx numbered from the input packet's sample count, like the reference)."""
        rsp = ray_samples_packed
        se, n_rays = VolumeRendering._prep(rsp, "importance_sample")
        n_imp = int(nr_importance_samples)
        S = rsp.get_max_nr_samples()
        if n_rays == 0 or S == 0:
            raise RuntimeError("RaySamplesPacked should not be empty")
        cdf = _f32c(samples_cdf, "samples_cdf", 1)
        if cdf.shape[0] != S:
            raise RuntimeError(f"CDF should have size of nr_samples_total x 1. but it has {tuple(cdf.shape)}")
        if n_imp <= 0:
            raise RuntimeError("nr_importance_samples must be positive")
        imp = RaySamplesPacked(n_rays, n_rays * n_imp, S, rsp.get_values_dim())
        imp.is_compacted = False
        check(
            _lib.lib().vs_importance_sample(
                ptr(_f32c(rsp.ray_o, "ray_o", 3)), ptr(_f32c(rsp.ray_d, "ray_d", 3)), ptr(se), ptr(_f32c(rsp.samples_z, "samples_z", 1)),
                ptr(cdf), n_rays, S, n_imp, VolumeRendering._rng_state, VolumeRendering._rng_inc, int(bool(jitter_samples)),
                ptr(imp.samples_3d), ptr(imp.samples_dirs), ptr(imp.samples_z), ptr(imp.ray_start_end_idx), _stream(),
            ),
            "vs_importance_sample",
        )
        if jitter_samples:
            VolumeRendering._rng_advance()
        out = imp.compact_to_valid_samples()
        if out.get_max_nr_samples() <= 0:
            raise RuntimeError("nr_samples_imp should be > 0")
        return out

    @staticmethod
    def combine_ray_samples_packets(ray_samples_packed_1, ray_samples_packed_2, min_dist_between_samples):
        """src/VolumeRendering.cu:550-669: per-ray z-ordered merge of two compacted packets; a sample closer than
        ``min_dist_between_samples`` to the previously kept one is dropped.  Returns a compacted packet (has_dt False)."""
        a, b = ray_samples_packed_1, ray_samples_packed_2
        se1, n_rays = VolumeRendering._prep(a, "combine_ray_samples_packets")
        se2, n_rays2 = VolumeRendering._prep(b, "combine_ray_samples_packets")
        if a.has_samples_values != b.has_samples_values:
            raise RuntimeError("They are supposed to both has or not have samples values")
        if a.get_values_dim() != b.get_values_dim():
            raise RuntimeError("They are supposed to have the same values dims")
        if n_rays != n_rays2:
            raise RuntimeError("They are supposed to have the same number of rays")
        if a.is_empty() and b.is_empty():
            raise RuntimeError("Both ray_samples_packed are empty")
        if a.is_empty():
            return b
        if b.is_empty():
            return a
        L = _lib.lib()
        st = _stream()
        dev = se1.device
        vd = a.get_values_dim()
        n1, n2 = a.get_max_nr_samples(), b.get_max_nr_samples()
        comb = RaySamplesPacked(n_rays, n1 + n2, 0, vd)
        comb.ray_o, comb.ray_d = a.ray_o.clone(), a.ray_d.clone()
        comb.ray_enter, comb.ray_exit, comb.ray_max_dt = a.ray_enter.clone(), a.ray_exit.clone(), a.ray_max_dt.clone()
        comb.has_samples_values = a.has_samples_values
        comb.has_dt = False
        comb.is_compacted = False
        scratch = torch.empty(max(int(L.vs_pack_scratch_bytes(n_rays)), 8), dtype=torch.uint8, device=dev)
        out_start = torch.empty(n_rays, dtype=torch.int32, device=dev)
        total_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.vs_combine_offsets(ptr(se1), ptr(se2), n_rays, ptr(out_start), ptr(total_dev), ptr(scratch), st), "vs_combine_offsets")

        def f(t, name, cols):
            return _f32c(t, name, cols)

        i1, i2 = a.samples_idx.contiguous(), b.samples_idx.contiguous()
        check(
            L.vs_combine_merge(
                n_rays, float(min_dist_between_samples), vd,
                ptr(se1), ptr(i1), ptr(f(a.samples_3d, "samples_3d", 3)), ptr(f(a.samples_dirs, "samples_dirs", 3)),
                ptr(f(a.samples_z, "samples_z", 1)), ptr(f(a.samples_values, "samples_values", vd)),
                ptr(se2), ptr(i2), ptr(f(b.samples_3d, "samples_3d", 3)), ptr(f(b.samples_dirs, "samples_dirs", 3)),
                ptr(f(b.samples_z, "samples_z", 1)), ptr(f(b.samples_values, "samples_values", vd)),
                ptr(out_start), ptr(comb.samples_idx), ptr(comb.samples_3d), ptr(comb.samples_dirs), ptr(comb.samples_z),
                ptr(comb.samples_values), ptr(comb.ray_start_end_idx), st,
            ),
            "vs_combine_merge",
        )
        out = comb.compact_to_valid_samples()
        if out.get_max_nr_samples() <= 0:
            raise RuntimeError("total_nr_samples should be > 0")
        return out

    # ---- backward ops --------------------------------------------------------------------------------------------
    @staticmethod
    def cumprod_one_minus_alpha_to_transmittance_backward(
        grad_transmittance, grad_bg_transmittance, ray_samples_packed, alpha, transmittance, bg_transmittance, cumsumLV
    ):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "cumprod backward")
        x = _f32c(alpha, "alpha", 1)
        gbg = _f32c(grad_bg_transmittance, "grad_bg_transmittance", 1)
        bg = _f32c(bg_transmittance, "bg_transmittance", 1)
        cs = _f32c(cumsumLV, "cumsumLV", 1)
        dx = torch.zeros_like(x)
        check(
            _lib.lib().vs_cumprod_bwd(ptr(se), ptr(gbg), ptr(x), ptr(bg), ptr(cs), ptr(dx), n_rays, x.shape[0], _stream()),
            "vs_cumprod_bwd",
        )
        return dx

    @staticmethod
    def cumprod_backward_fused(grad_transmittance, grad_bg_transmittance, ray_samples_packed, alpha, transmittance, bg_transmittance):
        """LV product + reverse cumsum + division in one launch (extension; same result as the two-step path)."""
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "cumprod backward")
        x = _f32c(alpha, "alpha", 1)
        gT = _f32c(grad_transmittance, "grad_transmittance", 1)
        gbg = _f32c(grad_bg_transmittance, "grad_bg_transmittance", 1)
        T = _f32c(transmittance, "transmittance", 1)
        bg = _f32c(bg_transmittance, "bg_transmittance", 1)
        dx = torch.zeros_like(x)
        check(
            _lib.lib().vs_cumprod_bwd_fused(ptr(se), ptr(gT), ptr(gbg), ptr(x), ptr(T), ptr(bg), ptr(dx), n_rays, x.shape[0], _stream()),
            "vs_cumprod_bwd_fused",
        )
        return dx

    @staticmethod
    def _integrate_backward(grad_result, ray_samples_packed, values, weights, dim):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, f"integrate_with_weights_{dim}d_backward")
        g = _f32c(grad_result, "grad_result", dim)
        v = _f32c(values, "values", dim)
        w = _f32c(weights, "weights", 1)
        dv = torch.zeros_like(v)
        dw = torch.zeros_like(w)
        check(
            _lib.lib().vs_integrate_bwd(
                ptr(se), ptr(g), ptr(v), ptr(w), ptr(dv), ptr(dw), dim, n_rays, v.shape[0], int(VolumeRendering.reference_bugs), _stream()
            ),
            "vs_integrate_bwd",
        )
        return dv, dw

    @staticmethod
    def integrate_with_weights_1d_backward(grad_result, ray_samples_packed, values, weights, result):
        return VolumeRendering._integrate_backward(grad_result, ray_samples_packed, values, weights, 1)

    @staticmethod
    def integrate_with_weights_3d_backward(grad_result, ray_samples_packed, values, weights, result):
        return VolumeRendering._integrate_backward(grad_result, ray_samples_packed, values, weights, 3)

    @staticmethod
    def sum_over_rays_backward(grad_values_sum_per_ray, grad_values_sum_per_sample, ray_samples_packed, sample_values):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "sum_over_rays_backward")
        v = _f32c(sample_values, "sample_values")
        d = v.shape[1]
        if d not in (1, 2, 3):
            raise RuntimeError("sum_over_rays_backward supports 1, 2 or 3 values per sample")
        gr = _f32c(grad_values_sum_per_ray, "grad_values_sum_per_ray", d)
        gs = _f32c(grad_values_sum_per_sample, "grad_values_sum_per_sample", d)
        dv = torch.zeros_like(v)
        check(_lib.lib().vs_sum_bwd(ptr(se), ptr(gr), ptr(gs), ptr(dv), d, n_rays, v.shape[0], _stream()), "vs_sum_bwd")
        return dv

    # ---- fused compositing (extension; the product's fast path) -------------------------------------------------
    @staticmethod
    def composite(ray_samples_packed, alpha, rgb, z=None, return_weights=False, mode=0):
        """rgb [N,3], depth [N,1], acc [N,1], bgT [N,1] (+ weights, transmittance [S,1]) in one launch."""
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "composite")
        a = _f32c(alpha, "alpha", 1)
        c = _f32c(rgb, "rgb", 3)
        zz = _f32c(ray_samples_packed.samples_z if z is None else z, "z", 1)
        VolumeRendering._rows(ray_samples_packed, a, c, zz)
        S = a.shape[0]
        if c.shape[0] != S or zz.shape[0] != S:
            raise RuntimeError("alpha, rgb and z must have the same number of samples")
        f = dict(dtype=torch.float32, device=a.device)
        out_rgb = torch.empty((n_rays, 3), **f)
        out_depth = torch.empty((n_rays, 1), **f)
        out_acc = torch.empty((n_rays, 1), **f)
        out_bgT = torch.empty((n_rays, 1), **f)
        w = torch.zeros((S, 1), **f) if return_weights else None
        T = torch.zeros((S, 1), **f) if return_weights else None
        check(
            _lib.lib().vs_composite_fwd(
                ptr(se), ptr(a), ptr(c), ptr(zz), ptr(out_rgb), ptr(out_depth), ptr(out_acc), ptr(out_bgT), ptr(w), ptr(T),
                n_rays, S, int(mode), _stream(),
            ),
            "vs_composite_fwd",
        )
        if return_weights:
            return out_rgb, out_depth, out_acc, out_bgT, w, T
        return out_rgb, out_depth, out_acc, out_bgT

    @staticmethod
    def composite_backward(ray_samples_packed, alpha, rgb, z, g_rgb, g_depth, g_acc, g_bgT, need_dz=False, mode=0):
        se, n_rays = VolumeRendering._prep(ray_samples_packed, "composite_backward")
        a = _f32c(alpha, "alpha", 1)
        c = _f32c(rgb, "rgb", 3)
        zz = _f32c(z, "z", 1)
        VolumeRendering._rows(ray_samples_packed, a, c, zz)
        S = a.shape[0]
        gC = _f32c(g_rgb, "g_rgb", 3)
        gD = _f32c(g_depth, "g_depth", 1)
        gA = _f32c(g_acc, "g_acc", 1)
        gB = _f32c(g_bgT, "g_bgT", 1)
        for name, t in (("g_rgb", gC), ("g_depth", gD), ("g_acc", gA), ("g_bgT", gB)):
            if t.shape[0] != n_rays:
                raise RuntimeError(f"{name} must have one row per ray")
        d_alpha = torch.empty_like(a)
        d_rgb = torch.empty_like(c)
        d_z = torch.empty_like(zz) if need_dz else None
        check(
            _lib.lib().vs_composite_bwd(
                ptr(se), ptr(a), ptr(c), ptr(zz), ptr(gC), ptr(gD), ptr(gA), ptr(gB), ptr(d_alpha), ptr(d_rgb), ptr(d_z),
                n_rays, S, int(mode), _stream(),
            ),
            "vs_composite_bwd",
        )
        return d_alpha, d_rgb, d_z


def _pcg_advance(state: int, inc: int, delta: int = 1 << 32) -> int:
    """pcg32::advance (kernels/volsurfs/pcg32.h:158-180) on a host copy of the generator"""
    m64 = (1 << 64) - 1
    cur_mult, cur_plus, acc_mult, acc_plus = 0x5851F42D4C957F2D, inc, 1, 0
    delta &= m64
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & m64
            acc_plus = (acc_plus * cur_mult + cur_plus) & m64
        cur_plus = ((cur_mult + 1) * cur_plus) & m64
        cur_mult = (cur_mult * cur_mult) & m64
        delta >>= 1
    return (acc_mult * state + acc_plus) & m64


def _rays_args(rays_o, rays_d, t_a, t_b=None):
    o, d = _f32c(rays_o, "rays_o", 3), _f32c(rays_d, "rays_d", 3)
    a = _f32c(t_a, "ray_t_entry", 1)
    b = None if t_b is None else _f32c(t_b, "ray_t_exit", 1)
    n = int(o.shape[0])
    if d.shape[0] != n or a.shape[0] != n or (b is not None and b.shape[0] != n):
        raise RuntimeError("rays_o, rays_d and the ray t tensors must have one row per ray")
    return o, d, a, b, n


class OccupancyGrid:
    """Occupancy grid in Morton order (include/volsurfs/OccupancyGrid.cuh:9-68, bound at PyBridge.cxx:33-68): the container, the two
    queries the samplers rest on, and the density-grid maintenance of training (voxel sample points, update_grid_values,
    update_grid_occupancy_with_density_values / _with_sdf_values, init_sphere_roi). (get_first_rays_sample_start_of_grid_occupied_regions, advance_ray_sample_to_next_occupied_voxel)."""

    #: host copy of the reference's static ``pcg32 m_rng`` (src/OccupancyGrid.cu:19)
    _rng_state = 0x853C49E6748FEA9B
    _rng_inc = 0xDA3E39CB94B95BDB

    def __init__(self, nr_voxels_per_dim: int, grid_extent):
        self.m_nr_voxels_per_dim = int(nr_voxels_per_dim)
        ext = [float(v) for v in (grid_extent.tolist() if hasattr(grid_extent, "tolist") else grid_extent)]
        if len(ext) != 3:
            raise ValueError("grid_extent must have 3 entries")
        self.m_grid_extent = ext
        self.m_grid_values = OccupancyGrid.make_grid_values(nr_voxels_per_dim)
        self.m_grid_occupancy = OccupancyGrid.make_grid_occupancy(nr_voxels_per_dim)
        self.m_grid_roi = OccupancyGrid.make_grid_occupancy(nr_voxels_per_dim)

    @staticmethod
    def _check_n(n: int) -> int:
        n = int(n)
        if n <= 0 or n % 2 != 0 or (n & (n - 1)) != 0:  # src/OccupancyGrid.cu:166-167 (CHECK-abort in the reference)
            raise RuntimeError("Nr of voxels should be an even power of 2 because we are using morton codes")
        return n

    @staticmethod
    def make_grid_values(nr_voxels_per_dim: int) -> torch.Tensor:
        n = OccupancyGrid._check_n(nr_voxels_per_dim)
        return torch.ones(n * n * n, dtype=torch.float32, device=_device())

    @staticmethod
    def make_grid_occupancy(nr_voxels_per_dim: int) -> torch.Tensor:
        n = OccupancyGrid._check_n(nr_voxels_per_dim)
        return torch.ones(n * n * n, dtype=torch.bool, device=_device())

    def get_grid_values(self):
        return self.m_grid_values

    def get_grid_occupancy(self):
        return self.m_grid_occupancy

    def get_grid_roi(self):
        return self.m_grid_roi

    def get_grid_occupancy_in_roi(self):
        return self.m_grid_occupancy.masked_select(self.m_grid_roi)

    def set_grid_values(self, grid_values):
        self.m_grid_values = grid_values

    def set_grid_occupancy(self, grid_occupancy):
        self.m_grid_occupancy = grid_occupancy

    def set_grid_roi(self, grid_roi):
        """extension: the reference only sets the region of interest through init_sphere_roi"""
        self.m_grid_roi = grid_roi

    def set_grid_occupancy_full(self):
        self.m_grid_occupancy.fill_(True)

    def set_grid_occupancy_empty(self):
        self.m_grid_occupancy.fill_(False)

    def get_nr_voxels(self) -> int:
        return self.m_nr_voxels_per_dim ** 3

    def get_nr_voxels_per_dim(self) -> int:
        return self.m_nr_voxels_per_dim

    def get_grid_extent(self):
        return list(self.m_grid_extent)

    def get_nr_voxels_in_roi(self) -> int:
        return int(self.m_grid_roi.sum().item())

    def get_nr_occupied_voxels(self) -> int:
        return int(self.m_grid_occupancy.sum().item())

    def get_nr_occupied_voxels_in_roi(self) -> int:
        return int(self.get_grid_occupancy_in_roi().sum().item())

    def get_grid_max_value(self) -> float:
        return float(self.m_grid_values.max().item())

    def get_grid_min_value(self) -> float:
        return float(self.m_grid_values.min().item())

    def get_grid_max_value_in_roi(self) -> float:
        return float(self.m_grid_values.masked_select(self.m_grid_roi).max().item())

    def get_grid_min_value_in_roi(self) -> float:
        return float(self.m_grid_values.masked_select(self.m_grid_roi).min().item())

    # ---- voxel sample points (src/OccupancyGrid.cu:206-347) -----------------------------------------------------------------------
    def _points(self, point_indices, centre: bool, jitter: bool):
        idx = point_indices
        if idx.dtype != torch.int32 or idx.dim() != 1 or not idx.is_cuda:
            raise RuntimeError("point_indices must be a 1-d int32 CUDA tensor")
        idx = idx.contiguous()
        n = int(idx.shape[0])
        out = torch.empty((n, 3), dtype=torch.float32, device=idx.device)
        check(_lib.lib().vs_occgrid_points(ptr(idx), self.m_nr_voxels_per_dim, self._extent_c(), int(centre), OccupancyGrid._rng_state,
                                           OccupancyGrid._rng_inc, int(bool(jitter)), ptr(out), n, _stream()), "vs_occgrid_points")
        if centre and jitter:
            OccupancyGrid._rng_state = _pcg_advance(OccupancyGrid._rng_state, OccupancyGrid._rng_inc)
        return out, idx

    def get_grid_lower_left_voxels_vertices(self):
        """(lower-left corner of every voxel [V,3], voxel indices [V] int32), Morton order (src/OccupancyGrid.cu:206-234)"""
        return self._points(torch.arange(0, self.get_nr_voxels(), dtype=torch.int32, device=_device()), False, False)

    def get_grid_samples(self, jitter_samples):
        """(centre of every voxel, optionally jittered inside it [V,3], voxel indices [V]) (src/OccupancyGrid.cu:236-271)"""
        return self._points(torch.arange(0, self.get_nr_voxels(), dtype=torch.int32, device=_device()), True, jitter_samples)

    def get_random_grid_samples(self, nr_voxels_to_select, jitter_samples):
        """nr_voxels_to_select voxels drawn with replacement (torch.randint, as the reference) (src/OccupancyGrid.cu:273-308)"""
        idx = torch.randint(0, self.get_nr_voxels(), (int(nr_voxels_to_select),), dtype=torch.int32, device=_device())
        return self._points(idx, True, jitter_samples)

    def get_random_grid_samples_in_roi(self, nr_voxels_to_select, jitter_samples):
        """the same, among the voxels of the region of interest (src/OccupancyGrid.cu:310-347)"""
        roi_idx = torch.nonzero(self.m_grid_roi).to(torch.int32)
        pick = torch.randint(0, int(roi_idx.shape[0]), (int(nr_voxels_to_select),), dtype=torch.int32, device=roi_idx.device)
        return self._points(roi_idx.index_select(0, pick).squeeze(1), True, jitter_samples)

    def init_sphere_roi(self, radius, padding):
        """region of interest = voxels whose 8 corners lie inside the sphere of radius - padding (src/OccupancyGrid.cu:117-151, the
        reference's torch expressions)"""
        ll, _ = self.get_grid_lower_left_voxels_vertices()
        n = float(self.m_nr_voxels_per_dim)
        voxel = torch.tensor(self.m_grid_extent, dtype=torch.float32, device=ll.device) / n
        offs = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=torch.float32,
                            device=ll.device) * voxel
        corners = (ll.view(-1, 1, 3) + offs.view(1, -1, 3)).reshape(-1, 3)
        inside = corners.norm(2, 1, True) < (float(radius) - float(padding))
        self.m_grid_roi = inside.reshape(-1, 8).all(-1)

    # ---- density-grid maintenance (src/OccupancyGrid.cu:446-503) ---------------------------------------------------------------------
    def update_grid_values(self, point_indices, values, decay):
        """grid_values[point_indices[i]] = max(values[i], decay * grid_values[point_indices[i]])"""
        if values.dim() != 2:
            raise RuntimeError(f"values should have shape nr_pointsx1. However it has sizes {tuple(values.shape)}")
        if float(decay) > 1.0:
            raise RuntimeError(f"We except the decay to be < 1.0 but it is {decay}")
        if point_indices.dim() != 1:
            raise RuntimeError(f"point_indices should have dim 1 correspondin to nr_points. However it has sizes {tuple(point_indices.shape)}")
        idx, v = self._indices(point_indices), _f32c(values, "values", 1)
        if v.shape[0] != idx.shape[0]:
            raise RuntimeError("values and point_indices must have the same number of rows")
        g = self._values_inplace()
        check(_lib.lib().vs_occgrid_update_values(ptr(idx), ptr(v), float(decay), ptr(g), int(idx.shape[0]), _stream()), "vs_occgrid_update_values")

    def update_grid_occupancy_with_density_values(self, point_indices, occupancy_tresh, check_neighbours):
        """occupancy[point_indices[i]] = value (or any value of the 3x3x3 neighbourhood) > occupancy_tresh"""
        if point_indices.dim() != 1:
            raise RuntimeError(f"point_indices should have dim 1 correspondin to nr_points. However it has sizes {tuple(point_indices.shape)}")
        idx = self._indices(point_indices)
        g = self._values_inplace()
        occ = self.m_grid_occupancy
        if occ.dtype != torch.bool or occ.numel() != self.get_nr_voxels() or not occ.is_cuda or not occ.is_contiguous():
            raise RuntimeError("grid_occupancy must be a contiguous bool CUDA tensor with nr_voxels_per_dim^3 entries")
        check(_lib.lib().vs_occgrid_update_occupancy_density(ptr(idx), self.m_nr_voxels_per_dim, self._extent_c(), float(occupancy_tresh),
                                                             int(bool(check_neighbours)), ptr(g), ptr(occ), int(idx.shape[0]), _stream()),
              "vs_occgrid_update_occupancy_density")

    def update_grid_occupancy_with_sdf_values(self, point_indices, logistic_beta, occupancy_thresh, check_neighbours):
        """occupancy[point_indices[i]] = logistic density (beta_i) at the closest the surface can be inside the voxel > occupancy_thresh
        (src/OccupancyGrid.cu:505-533; check_neighbours is accepted and unused, as in the reference's kernel)"""
        if point_indices.dim() != 1:
            raise RuntimeError(f"point_indices should have dim 1 correspondin to nr_points. However it has sizes {tuple(point_indices.shape)}")
        idx = self._indices(point_indices)
        beta = _f32c(logistic_beta, "logistic_beta", 1)
        if beta.shape[0] != idx.shape[0]:
            raise RuntimeError("logistic_beta must have one row per point index")
        g = self._values_inplace()
        occ = self.m_grid_occupancy
        if occ.dtype != torch.bool or occ.numel() != self.get_nr_voxels() or not occ.is_cuda or not occ.is_contiguous():
            raise RuntimeError("grid_occupancy must be a contiguous bool CUDA tensor with nr_voxels_per_dim^3 entries")
        check(_lib.lib().vs_occgrid_update_occupancy_sdf(ptr(idx), self.m_nr_voxels_per_dim, self._extent_c(), ptr(beta), float(occupancy_thresh),
                                                         ptr(g), ptr(occ), int(idx.shape[0]), _stream()), "vs_occgrid_update_occupancy_sdf")

    def _indices(self, point_indices):
        if point_indices.dtype != torch.int32 or not point_indices.is_cuda:
            raise RuntimeError("point_indices must be an int32 CUDA tensor")
        return point_indices.contiguous()

    def _values_inplace(self):
        g = self.m_grid_values
        if g.dtype != torch.float32 or g.numel() != self.get_nr_voxels() or not g.is_cuda or not g.is_contiguous():
            raise RuntimeError("grid_values must be a contiguous float32 CUDA tensor with nr_voxels_per_dim^3 entries")
        return g

    def _masks(self):
        n3 = self.get_nr_voxels()
        occ, roi = self.m_grid_occupancy, self.m_grid_roi
        for name, t in (("grid_occupancy", occ), ("grid_roi", roi)):
            if t.dtype != torch.bool or t.numel() != n3 or not t.is_cuda:
                raise RuntimeError(f"{name} must be a bool CUDA tensor with nr_voxels_per_dim^3 entries")
        return occ.contiguous(), roi.contiguous()

    def _extent_c(self):
        import ctypes

        return (ctypes.c_float * 3)(*self.m_grid_extent)

    def get_rays_t_near_t_far(self, rays_o, rays_d, ray_t_entry, ray_t_exit):
        """first / last t inside occupied voxels of the region of interest (src/OccupancyGrid.cu:349-399) -> ([N,1], [N,1])"""
        o, d, a, b, n = _rays_args(rays_o, rays_d, ray_t_entry, ray_t_exit)
        occ, roi = self._masks()
        near = torch.empty((n, 1), dtype=torch.float32, device=o.device)
        far = torch.empty((n, 1), dtype=torch.float32, device=o.device)
        check(_lib.lib().vs_occgrid_rays_t_near_t_far(ptr(o), ptr(d), ptr(a), ptr(b), self.m_nr_voxels_per_dim, self._extent_c(), ptr(occ),
                                                      ptr(roi), ptr(near), ptr(far), n, _stream()), "vs_occgrid_rays_t_near_t_far")
        return near, far

    def get_first_rays_sample_start_of_grid_occupied_regions(self, rays_o, rays_d, ray_t_entry, ray_t_exit):
        """the sphere tracer's starting packet (src/OccupancyGrid.cu:536-573): RaySamplesPacked(nr_rays, nr_rays, 0, 1) holding, at row =
        ray index, the point where the ray first probes an occupied voxel of the roi; rays without one have the segment (0, 0)"""
        o, d, a, b, n = _rays_args(rays_o, rays_d, ray_t_entry, ray_t_exit)
        occ, roi = self._masks()
        out = RaySamplesPacked(n, n, 0, 1)
        check(_lib.lib().vs_occgrid_first_sample_start(ptr(o), ptr(d), ptr(a), ptr(b), self.m_nr_voxels_per_dim, self._extent_c(), ptr(occ),
                                                       ptr(roi), ptr(out.samples_3d), ptr(out.samples_dirs), ptr(out.samples_z),
                                                       ptr(out.samples_dt), ptr(out.ray_start_end_idx), n, _stream()),
              "vs_occgrid_first_sample_start")
        return out

    def advance_ray_sample_to_next_occupied_voxel(self, samples_dirs, samples_3d):
        """every point marched along its direction to the first occupied voxel of the roi, or to the last position inside the grid
        (is_within_bounds False; a point leaving through a LOWER face ends the march too — the reference's kernel does not return for those).
        As in the reference (src/OccupancyGrid.cu:575-607, ``new_samples_3d = samples_3d``) a contiguous
        float32 input is updated IN PLACE and returned -> (new_samples_3d [P,3], is_within_bounds [P,1] bool)"""
        p, d = _f32c(samples_3d, "samples_3d", 3), _f32c(samples_dirs, "samples_dirs", 3)
        if d.shape[0] != p.shape[0]:
            raise RuntimeError("samples_dirs and samples_3d must have the same number of rows")
        occ, roi = self._masks()
        n = int(p.shape[0])
        within = torch.ones((n, 1), dtype=torch.bool, device=p.device)
        check(_lib.lib().vs_occgrid_advance_to_next_occupied(ptr(d), ptr(p), self.m_nr_voxels_per_dim, self._extent_c(), ptr(occ), ptr(roi), ptr(p),
                                                             ptr(within), n, _stream()), "vs_occgrid_advance_to_next_occupied")
        return p, within

    def check_occupancy(self, points):
        """(occupied && in roi [P,1] bool, grid value [P,1]) per point; outside the grid -> (False, 0) (src/OccupancyGrid.cu:402-447)"""
        p = _f32c(points, "points", 3)
        occ, roi = self._masks()
        vals = self.m_grid_values
        if vals.dtype != torch.float32 or vals.numel() != self.get_nr_voxels():
            raise RuntimeError("grid_values must be a float32 tensor with nr_voxels_per_dim^3 entries")
        n = int(p.shape[0])
        out_occ = torch.empty((n, 1), dtype=torch.bool, device=p.device)
        out_val = torch.empty((n, 1), dtype=torch.float32, device=p.device)
        check(_lib.lib().vs_occgrid_check_occupancy(ptr(p), self.m_nr_voxels_per_dim, self._extent_c(), ptr(vals.contiguous()), ptr(occ), ptr(roi),
                                                    ptr(out_occ), ptr(out_val), n, _stream()), "vs_occgrid_check_occupancy")
        return out_occ, out_val


class RaySampler:
    """Static ray samplers (include/volsurfs/RaySampler.cuh:9-64, bound at PyBridge.cxx:131-139).  The foreground samplers return the
    COMPACTED packet, like the reference after its closing compact_to_valid_samples (src/RaySampler.cu:236,340), but never build the
    nr_rays x max_nr_samples_per_ray staging packet of 36-byte rows: only 4 bytes of depth per slot are staged (csrc/sampler.cu)."""

    #: host copy of the reference's static ``pcg32 m_rng``: passed by value to a jittered launch, advanced by 2^32 afterwards
    _rng_state = 0x853C49E6748FEA9B
    _rng_inc = 0xDA3E39CB94B95BDB

    @staticmethod
    def _fg(rays_o, rays_d, ray_t_entry, ray_t_exit, min_dist, min_nr, max_nr, jitter, values_dim, grid):
        import ctypes

        L = _lib.lib()
        o, d, a, b, n = _rays_args(rays_o, rays_d, ray_t_entry, ray_t_exit)
        dev, st = o.device, _stream()
        min_nr, max_nr = int(min_nr), int(max_nr)
        if n * max_nr > 2**31 - 1:
            raise RuntimeError("nr_rays * max_nr_samples_per_ray must fit int32 sample indices")
        if grid is None:
            nv, ext, occ, roi = 0, None, None, None
        else:
            nv, ext_list, occ, roi = grid
            ext = (ctypes.c_float * 3)(*[float(v) for v in ext_list])
        se_virtual = torch.empty((n, 2), dtype=torch.int32, device=dev)
        ray_max_dt = torch.full((n, 1), -1.0, dtype=torch.float32, device=dev)
        z_stage = torch.empty(max(n * max_nr, 1), dtype=torch.float32, device=dev)  # 4 B per slot; only real samples are touched
        check(L.vs_sampler_fg_count(ptr(o), ptr(d), ptr(a), ptr(b), float(min_dist), min_nr, max_nr, RaySampler._rng_state, RaySampler._rng_inc,
                                    int(bool(jitter)), int(nv), ext, ptr(occ), ptr(roi), ptr(se_virtual), ptr(ray_max_dt), ptr(z_stage), n, st),
              "vs_sampler_fg_count")
        scratch = torch.empty(max(int(L.vs_pack_scratch_bytes(n)), 8), dtype=torch.uint8, device=dev)
        out_start = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        total_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        check(L.vs_segment_offsets(ptr(se_virtual), n, ptr(out_start), ptr(total_dev), ptr(scratch), st), "vs_segment_offsets")
        total = int(total_dev.item())  # sizes the outputs exactly, like the reference's compaction
        f = dict(dtype=torch.float32, device=dev)
        out = RaySamplesPacked._from_tensors(
            samples_idx=torch.empty((total, 1), dtype=torch.int32, device=dev),
            samples_3d=torch.empty((total, 3), **f),
            samples_dirs=torch.empty((total, 3), **f),
            samples_z=torch.empty((total, 1), **f),
            samples_dt=torch.full((total, 1), -1.0, **f),                 # the samplers never write dt (update_dt does)
            samples_values=torch.full((total, int(values_dim)), -1.0, **f),
            ray_start_end_idx=torch.empty((n, 2), dtype=torch.int32, device=dev),
            ray_o=o.clone(), ray_d=d.clone(), ray_enter=a.clone(), ray_exit=b.clone(), ray_max_dt=ray_max_dt,
        )
        check(L.vs_sampler_fg_write(ptr(o), ptr(d), max_nr, ptr(se_virtual), ptr(z_stage), ptr(out_start), ptr(out.ray_start_end_idx),
                                    ptr(out.samples_idx), ptr(out.samples_3d), ptr(out.samples_dirs), ptr(out.samples_z), n, st),
              "vs_sampler_fg_write")
        if jitter:
            RaySampler._rng_state = _pcg_advance(RaySampler._rng_state, RaySampler._rng_inc)
        return out

    @staticmethod
    def compute_samples_fg(rays_o, rays_d, ray_t_entry, ray_t_exit, min_dist_between_samples, min_nr_samples_per_ray, max_nr_samples_per_ray,
                           jitter_samples, values_dim):
        """equidistant samples between ray entry and exit (src/RaySampler.cu:159-245)"""
        return RaySampler._fg(rays_o, rays_d, ray_t_entry, ray_t_exit, min_dist_between_samples, min_nr_samples_per_ray,
                              max_nr_samples_per_ray, jitter_samples, values_dim, None)

    @staticmethod
    def compute_samples_fg_in_grid_occupied_regions(rays_o, rays_d, ray_t_entry, ray_t_exit, min_dist_between_samples, min_nr_samples_per_ray,
                                                    max_nr_samples_per_ray, jitter_samples, nr_voxels_per_dim, grid_extent, grid_occupancy,
                                                    grid_roi, values_dim):
        """samples spread over the occupied voxels a ray crosses (src/RaySampler.cu:247-345)"""
        n3 = int(nr_voxels_per_dim) ** 3
        for name, t in (("grid_occupancy", grid_occupancy), ("grid_roi", grid_roi)):
            if t.dtype != torch.bool or t.numel() != n3 or not t.is_cuda:
                raise RuntimeError(f"{name} must be a bool CUDA tensor with nr_voxels_per_dim^3 entries")
        ext = grid_extent.tolist() if hasattr(grid_extent, "tolist") else list(grid_extent)
        return RaySampler._fg(rays_o, rays_d, ray_t_entry, ray_t_exit, min_dist_between_samples, min_nr_samples_per_ray,
                              max_nr_samples_per_ray, jitter_samples, values_dim,
                              (int(nr_voxels_per_dim), ext, torch.logical_and(grid_occupancy, grid_roi).contiguous(), None))

    @staticmethod
    def compute_samples_bg(rays_o, rays_d, ray_t_exit, ray_t_far, nr_samples, jitter_samples):
        """``nr_samples`` samples per ray from the foreground exit to ``ray_t_far``, uniform in inverse depth (src/RaySampler.cu:72-157)"""
        o, d, a, _, n = _rays_args(rays_o, rays_d, ray_t_exit)
        nr = int(nr_samples)
        out = RaySamplesPacked(n, n * nr, 0, 0)
        out.ray_o, out.ray_d, out.ray_enter = o.clone(), d.clone(), a.clone()
        out.ray_exit = torch.full((n, 1), float(ray_t_far), dtype=torch.float32, device=o.device)
        out.is_compacted = True
        check(_lib.lib().vs_sampler_bg(ptr(o), ptr(d), ptr(a), float(ray_t_far), nr, RaySampler._rng_state, RaySampler._rng_inc,
                                       int(bool(jitter_samples)), ptr(out.ray_max_dt), ptr(out.samples_3d), ptr(out.samples_dirs),
                                       ptr(out.samples_z), ptr(out.ray_start_end_idx), n, _stream()), "vs_sampler_bg")
        if jitter_samples:
            RaySampler._rng_state = _pcg_advance(RaySampler._rng_state, RaySampler._rng_inc)
        return out

    @staticmethod
    def init_with_one_sample_per_ray(samples_3d, samples_dirs):
        """one sample per ray at z = 0 (src/RaySampler.cu:28-70; kernel RaySamplerGPU.cuh:490-526): a packet of copies, built from device
        tensor copies — there is no arithmetic to put in a kernel"""
        p, d = _f32c(samples_3d, "samples_3d", 3), _f32c(samples_dirs, "samples_dirs", 3)
        n = int(p.shape[0])
        out = RaySamplesPacked(n, n, 0, 1)
        out.samples_3d, out.samples_dirs = p.clone(), d.clone()
        out.samples_z.zero_()
        out.samples_dt.zero_()
        idx = torch.arange(n, dtype=torch.int32, device=p.device)
        out.ray_start_end_idx = torch.stack([idx, idx + 1], dim=1).contiguous()
        return out

    @staticmethod
    def _contract(rsp, uncontract: bool, what: str):
        if rsp.is_empty():
            raise RuntimeError(f"RaySamplesPacked must not be empty before calling {what}")
        if not rsp.is_compacted:
            raise RuntimeError(f"RaySamplesPacked should be compacted at the beginning of {what}")
        out = rsp.copy()
        n = rsp.get_nr_rays()
        check(_lib.lib().vs_sampler_contract(ptr(_f32c(rsp.ray_o, "ray_o", 3)), ptr(rsp.ray_start_end_idx.contiguous()),
                                             ptr(_f32c(rsp.samples_3d, "samples_3d", 3)), ptr(_f32c(rsp.samples_z, "samples_z", 1)),
                                             ptr(out.samples_3d), ptr(out.samples_z), int(uncontract), n, _stream()), "vs_sampler_contract")
        out.update_dt(True)  # src/RaySampler.cu:378,425
        return out

    @staticmethod
    def contract_samples(uncontracted_ray_samples_packed):
        """background scene contraction (src/RaySampler.cu:336-381): positions beyond |2x| = 1 are pulled into the unit ball, depths
        re-measured from the ray origin, dt updated as for a background packet"""
        return RaySampler._contract(uncontracted_ray_samples_packed, False, "contract_samples")

    @staticmethod
    def uncontract_samples(contracted_ray_samples_packed):
        """inverse of contract_samples (src/RaySampler.cu:383-427)"""
        return RaySampler._contract(contracted_ray_samples_packed, True, "uncontract_samples")
