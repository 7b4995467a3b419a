"""Permutohedral-lattice hash encoding: host-side mirror of the reference's interface for this stage.

* :class:`PermutoEncoding`, :class:`Coarse2Fine`, :func:`cosine_easing_window` mirror the package ``permutohedral_encoding``
  (submodules/permutohedral_encoding/src/pytorch_modules/modules.py:11-120, funcs.py:8-76, utils.py:5-22): same constructor
  arguments, parameter names and shapes (``lattice_values`` [levels, capacity, 2], ``random_shift_per_level`` [levels, pos_dim]),
  ``forward(positions, anneal_window)`` -> [N, output_dims()], ``output_dims()``, ``reset()``.
* :class:`PermutoHashEncoder` mirrors volsurfs_py/encodings/permutohash.py:10-99 (the encoder the legacy ``RGB`` heads own):
  ``__call__(points, iter_nr)`` -> (``enc_points`` [N, output_dim], ``points_out_of_bounds`` [N] bool).

The kernels (csrc/permuto.cu) write complete rows — the operand layout of :class:`~volsurfs_b200.appearance.AppearanceHead` — and
fold the bounding-box normalisation, the out-of-bounds mask and ``remove_last_element`` into the same launch.  No CPU fallback.
``install_as_permutohedral_encoding()`` registers this module under the reference's package name."""
from __future__ import annotations

import ctypes
import math
import sys

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .volsurfs import _stream


def cosine_easing_window(num_freqs: int, alpha) -> torch.Tensor:
    """utils.py:5-17"""
    x = torch.clip(alpha - torch.arange(num_freqs, dtype=torch.float32), 0.0, 1.0)
    return 0.5 * (1 + torch.cos(math.pi * x + math.pi))


def map_range_val(input_val, input_start, input_end, output_start, output_end):
    """utils.py:19-22 / volsurfs_py/utils/common.py"""
    if input_end == input_start:   # the reference divides by zero here (nr_iters_for_c2f = 0); any iter_nr >= 0 then means "fully open"
        return output_end if input_val >= input_end else output_start
    clamped = max(input_start, min(input_end, input_val))
    return output_start + ((output_end - output_start) / (input_end - input_start)) * (clamped - input_start)


def scale_factors(scale_per_level, pos_dim: int) -> torch.Tensor:
    """Encoding.cuh:53-67: 1/sqrt((i+1)(i+2)) rounded to fp32, times the fp32 reciprocal of sigma (torch divides a CUDA tensor by a
    host scalar as a multiplication with the reciprocal).  Host tensor [levels, pos_dim]."""
    sig = np.asarray(scale_per_level, dtype=np.float64).astype(np.float32)
    out = np.zeros((len(sig), pos_dim), np.float32)
    for r in range(len(sig)):
        inv = np.float32(1.0) / sig[r]
        for i in range(pos_dim):
            out[r, i] = np.float32(1.0 / math.sqrt(float((i + 1) * (i + 2)))) * inv
    return torch.from_numpy(out)


class _PermutoFunction(torch.autograd.Function):
    """PermutoEncodingFunc / PermutoEncodingFuncBack (funcs.py:8-55) in one node; no double backward (the volsurfs path never asks
    for second derivatives of the appearance encoders)."""

    @staticmethod
    def forward(ctx, enc, lattice_values, positions, window, out_cols, bb_sides, n_valid_dev):
        out, oob = enc._launch_forward(lattice_values, positions, window, out_cols, bb_sides, n_valid_dev)
        ctx.enc, ctx.window, ctx.bb_sides, ctx.n_valid_dev = enc, window, bb_sides, n_valid_dev
        ctx.save_for_backward(lattice_values, positions)
        ctx.mark_non_differentiable(oob)
        return out, oob

    @staticmethod
    def backward(ctx, grad_out, _grad_oob):
        lattice_values, positions = ctx.saved_tensors
        if grad_out is None:  # only the out-of-bounds mask was used downstream
            return (None, torch.zeros_like(lattice_values) if ctx.needs_input_grad[1] else None,
                    torch.zeros_like(positions) if ctx.needs_input_grad[2] else None, None, None, None, None)
        grad_out = grad_out.contiguous()
        d_lat, d_pos = ctx.enc._launch_backward(lattice_values, positions, ctx.window, grad_out, ctx.bb_sides, ctx.n_valid_dev,
                                                want_lattice=ctx.needs_input_grad[1], want_positions=ctx.needs_input_grad[2])
        return None, d_lat, d_pos, None, None, None, None


class PermutoEncoding(torch.nn.Module):
    def __init__(self, pos_dim, capacity, nr_levels, nr_feat_per_level, scale_per_level, appply_random_shift_per_level=True,
                 concat_points=False, concat_points_scaling=1.0, device=None):
        super().__init__()
        if nr_feat_per_level != 2:
            raise RuntimeError("Encoding: nr_feat_per_level must be 2 since other values are not yet implemented")   # Encoding.cuh:153-156
        if not 2 <= pos_dim <= 4:
            raise RuntimeError("Encoding: pos_dim must be 2, 3 or 4 in this build (the reference instantiates 2..7)")
        assert len(scale_per_level) == nr_levels
        self.pos_dim, self.capacity, self.nr_levels, self.nr_feat_per_level = int(pos_dim), int(capacity), int(nr_levels), 2
        self.scale_per_level = scale_per_level
        self.appply_random_shift_per_level = appply_random_shift_per_level
        self.concat_points, self.concat_points_scaling = bool(concat_points), float(concat_points_scaling)
        self._device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else torch.device("cpu")
        self.reset()
        if appply_random_shift_per_level:
            shift = torch.randn(nr_levels, pos_dim) * 10
        else:
            shift = torch.zeros(nr_levels, pos_dim)   # the reference allocates an (unused-shape) placeholder; zeros = no shift
        self.random_shift_per_level = torch.nn.Parameter(shift.to(self._device))   # a Parameter so that checkpoints carry it (modules.py:29)
        self.register_buffer("anneal_window", torch.ones(nr_levels, device=self._device), persistent=False)
        self.register_buffer("scale_factor", scale_factors(scale_per_level, pos_dim).to(self._device), persistent=False)

    def reset(self):
        """modules.py:40-44"""
        lattice_values = torch.randn(self.capacity, self.nr_levels, self.nr_feat_per_level) * 1e-5
        self.lattice_values = torch.nn.Parameter(lattice_values.permute(1, 0, 2).contiguous().to(self._device))

    def output_dims(self) -> int:
        extra = math.ceil(float(self.pos_dim) / self.nr_feat_per_level) if self.concat_points else 0
        return self.nr_feat_per_level * (self.nr_levels + extra)

    # ---- launches ------------------------------------------------------------------------------------------------------------
    def _check(self, positions):
        if not positions.is_cuda:
            raise _lib.VolsurfsB200Error("positions should be in GPU memory (there is no CPU path for the permutohedral encoding)")
        if positions.dtype != torch.float32 or positions.dim() != 2 or positions.shape[1] != self.pos_dim:
            raise ValueError(f"positions should be float32 [N, {self.pos_dim}], got {positions.dtype} {tuple(positions.shape)}")

    @staticmethod
    def _bb_c(bb_sides, pos_dim):
        if bb_sides is None:
            return None
        return (ctypes.c_float * pos_dim)(*[float(b) for b in bb_sides])

    def _launch_forward(self, lattice_values, positions, window, out_cols, bb_sides, n_valid_dev, out=None):
        self._check(positions)
        positions = positions.contiguous()
        n = int(positions.shape[0])
        if out is None:
            out = torch.empty((n, out_cols), dtype=torch.float32, device=positions.device)
        oob = torch.zeros((n,), dtype=torch.bool, device=positions.device) if bb_sides is not None else torch.empty(0, dtype=torch.bool, device=positions.device)
        check(_lib.lib().vs_permuto_forward(
            self.pos_dim, self.nr_levels, self.capacity, int(self.concat_points), self.concat_points_scaling, self._bb_c(bb_sides, self.pos_dim),
            ptr(positions), ptr(lattice_values.detach()), ptr(self.scale_factor), ptr(self.random_shift_per_level.detach()),
            ptr(window), ptr(out), int(out_cols), int(out.stride(0)), ptr(oob) if bb_sides is not None else None, n, ptr(n_valid_dev), _stream()),
            "vs_permuto_forward")
        return out, oob

    def _launch_backward(self, lattice_values, positions, window, grad_out, bb_sides, n_valid_dev, want_lattice=True, want_positions=False,
                         d_lattice=None, d_positions=None, levels=None, order_key=None):
        """``levels=(l0, l1)``: lattice gradient of that level range only (the levels are independent: every argument of the C entry
        point is simply offset to the range), so that a caller can hand finished level blocks to the gradient exchange while the next block
        is still being computed.  ``order_key`` [N] int32 (e.g. ``RaySamplesPacked.samples_layer``): ordering hint, see
        ``vs_permuto_backward_keyed`` in include/volsurfs_b200.h"""
        positions = positions.contiguous()
        grad_out = grad_out.contiguous()
        n = int(positions.shape[0])
        if order_key is not None:
            order_key = order_key.view(-1)
            assert order_key.dtype == torch.int32 and order_key.is_contiguous() and order_key.shape[0] >= n
        if want_lattice and d_lattice is None:
            d_lattice = torch.zeros_like(lattice_values)
        if want_positions and d_positions is None:
            d_positions = torch.zeros_like(positions)
        if levels is not None:
            l0, l1 = int(levels[0]), int(levels[1])
            assert want_lattice and not want_positions and 0 <= l0 < l1 <= self.nr_levels
            cols = min(2 * (l1 - l0), int(grad_out.shape[1]) - 2 * l0)
            assert cols >= 1
            fs = 4  # bytes per float
            check(_lib.lib().vs_permuto_backward_keyed(
                self.pos_dim, l1 - l0, self.capacity, 0, self._bb_c(bb_sides, self.pos_dim), ptr(positions), ptr(order_key),
                lattice_values.data_ptr() + l0 * self.capacity * 2 * fs, self.scale_factor.data_ptr() + l0 * self.pos_dim * fs,
                self.random_shift_per_level.data_ptr() + l0 * self.pos_dim * fs, window.data_ptr() + l0 * fs,
                grad_out.data_ptr() + 2 * l0 * fs, cols, int(grad_out.stride(0)), d_lattice.data_ptr() + l0 * self.capacity * 2 * fs,
                None, n, ptr(n_valid_dev), _stream()), "vs_permuto_backward_keyed")
            return d_lattice, None
        check(_lib.lib().vs_permuto_backward_keyed(
            self.pos_dim, self.nr_levels, self.capacity, int(self.concat_points), self._bb_c(bb_sides, self.pos_dim), ptr(positions),
            ptr(order_key), ptr(lattice_values.detach()), ptr(self.scale_factor), ptr(self.random_shift_per_level.detach()), ptr(window), ptr(grad_out),
            int(grad_out.shape[1]), int(grad_out.stride(0)), ptr(d_lattice) if want_lattice else None,
            ptr(d_positions) if want_positions else None, n, ptr(n_valid_dev), _stream()), "vs_permuto_backward_keyed")
        return (d_lattice if want_lattice else None), (d_positions if want_positions else None)

    # ---- reference interface -----------------------------------------------------------------------------------------------------
    def forward(self, positions, anneal_window=None, out_cols=None, bb_sides=None, n_valid_dev=None, return_out_of_bounds=False):
        """positions [N, pos_dim] -> [N, output_dims()] (modules.py:63-92).  Extensions used by the fused path: ``out_cols`` keeps the
        first columns only, ``bb_sides`` normalises the points inside the kernel, ``n_valid_dev`` bounds the rows on the device."""
        window = self.anneal_window if anneal_window is None else anneal_window.to(positions.device, torch.float32).contiguous().view(-1)
        out_cols = self.output_dims() if out_cols is None else int(out_cols)
        needs_grad = torch.is_grad_enabled() and (self.lattice_values.requires_grad or positions.requires_grad)
        if needs_grad:
            out, oob = _PermutoFunction.apply(self, self.lattice_values, positions, window, out_cols, bb_sides, n_valid_dev)
        else:
            out, oob = self._launch_forward(self.lattice_values, positions, window, out_cols, bb_sides, n_valid_dev)
        return (out, oob) if return_out_of_bounds else out


class Coarse2Fine(torch.nn.Module):
    """modules.py:101-120"""

    def __init__(self, nr_values):
        super().__init__()
        self.nr_values = nr_values
        self.last_t = 0

    def forward(self, t):
        assert t <= 1.0, "t cannot be larger than 1.0"
        window = cosine_easing_window(self.nr_values, t * self.nr_values)
        self.last_t = t
        return window

    def get_last_t(self):
        return self.last_t


class PermutoHashEncoder:
    """volsurfs_py/encodings/permutohash.py:10-99.  ``__call__`` returns ``(enc_points, points_out_of_bounds)``; both come out of one
    kernel launch (the reference runs ~8 elementwise kernels, the encoder, a permute and a slice)."""

    def __init__(self, input_dim=3, nr_levels=24, log2_hashmap_size=18, nr_feat_per_level=2, coarsest_scale=1.0, finest_scale=0.0001,
                 nr_iters_for_c2f=0, appply_random_shift_per_level=True, concat_points=True, concat_points_scaling=1.0,
                 remove_last_element=True, bb_sides=2.0, device=None):
        capacity = pow(2, log2_hashmap_size)
        scale_list = np.geomspace(coarsest_scale, finest_scale, num=nr_levels)
        self.encoder = PermutoEncoding(input_dim, capacity, nr_levels, nr_feat_per_level, scale_list,
                                       appply_random_shift_per_level=appply_random_shift_per_level, concat_points=concat_points,
                                       concat_points_scaling=concat_points_scaling, device=device)
        self.remove_last_element = remove_last_element
        self.input_dim = input_dim
        self.output_dim = self.encoder.output_dims() - 1 if remove_last_element else self.encoder.output_dims()
        self.capacity, self.nr_levels, self.nr_feat_per_level, self.scale_list = capacity, nr_levels, nr_feat_per_level, scale_list
        self.appply_random_shift_per_level, self.concat_points, self.concat_points_scaling = appply_random_shift_per_level, concat_points, concat_points_scaling
        if bb_sides is not None:
            if isinstance(bb_sides, (float, int)):
                bb_sides = [float(bb_sides)] * input_dim
            if isinstance(bb_sides, torch.Tensor):       # the reference accepts tensors (and moves them to CUDA): permutohash.py:30-36
                bb_sides = bb_sides.detach().cpu().tolist()
            bb_sides = [float(b) for b in np.asarray(bb_sides, dtype=np.float32).reshape(-1)]
        self.bb_sides = bb_sides                     # host floats: they parameterise the kernel, no device tensor needed
        self.c2f = Coarse2Fine(nr_levels)
        self.nr_iters_for_c2f = nr_iters_for_c2f
        self._window_cache = {}

    def parameters(self):
        return self.encoder.parameters()

    def window(self, iter_nr=None):
        t = 1.0 if iter_nr is None or iter_nr < 0 else map_range_val(iter_nr, 0.0, self.nr_iters_for_c2f, 0.3, 1.0)
        w = self._window_cache.get(t)
        if w is None:
            w = self.c2f(t).view(-1).to(self.encoder.lattice_values.device)
            self._window_cache = {t: w}
        return w

    def __call__(self, points, iter_nr=None, n_valid_dev=None, **kwargs):
        enc_points, oob = self.encoder(points, self.window(iter_nr), out_cols=self.output_dim, bb_sides=self.bb_sides, n_valid_dev=n_valid_dev,
                                       return_out_of_bounds=True)
        return enc_points, (oob if self.bb_sides is not None else None)

    def reset(self):
        self.encoder.reset()


def install_as_permutohedral_encoding() -> None:
    """``import permutohedral_encoding as permuto_enc`` (permutohash.py:4) then resolves to this module."""
    sys.modules["permutohedral_encoding"] = sys.modules[__name__]
