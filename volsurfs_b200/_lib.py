"""ctypes binding of ``libvolsurfs_b200.so`` (C ABI declared in ``include/volsurfs_b200.h``).

There is no CPU fallback: if the shared object is missing or a symbol cannot be resolved the import of the
operators fails loudly.  ``lib()`` only *loads* the library (possible without a GPU); calling a kernel
entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libvolsurfs_b200.so"

_P = c_void_p  # every device pointer
_I64 = c_int64

# name -> (restype, argtypes).  Kept in the order of include/volsurfs_b200.h.
SIGNATURES = {
    "vs_abi_version": (c_int, []),
    "vs_error_string": (c_char_p, [c_int]),
    "vs_launch_count": (ctypes.c_longlong, []),
    "vs_cumprod_fwd": (c_int, [_P, _P, _P, _P, _I64, _I64, _P]),
    "vs_cumprod_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "vs_cumprod_bwd_fused": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "vs_cumsum": (c_int, [_P, _P, _P, c_int, _I64, _I64, _P]),
    "vs_integrate_fwd": (c_int, [_P, _P, _P, _P, c_int, _I64, _I64, _P]),
    "vs_integrate_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _I64, _I64, c_int, _P]),
    "vs_sum_fwd": (c_int, [_P, _P, _P, _P, c_int, _I64, _I64, _P]),
    "vs_sum_bwd": (c_int, [_P, _P, _P, _P, c_int, _I64, _I64, _P]),
    "vs_update_dt": (c_int, [_P, _P, _P, _P, _P, c_int, _I64, _I64, _P]),
    "vs_sdf2alpha": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, _P]),
    "vs_median_depth": (c_int, [_P, _P, _P, c_float, _P, _I64, _I64, c_int, _P]),
    "vs_compute_cdf": (c_int, [_P, _P, _P, _I64, _I64, _P]),
    "vs_composite_fwd": (c_int, [_P] * 10 + [_I64, _I64, c_int, _P]),
    "vs_composite_bwd": (c_int, [_P] * 11 + [_I64, _I64, c_int, _P]),
    "vs_pack_scratch_bytes": (_I64, [_I64]),
    "vs_count_total": (c_int, [_P, _I64, _P, _P]),
    "vs_compact_offsets": (c_int, [_P, _I64, _P, _P, _P]),
    "vs_compact_gather": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "vs_pack_hits_offsets": (c_int, [_P, c_int, c_float, _I64, _P, _P, _P]),
    "vs_pack_hits_scatter": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_float] + [_P] * 8 + [_I64, _P]),
    "vs_shells_build": (c_int, [c_int, _P, _P, _P, _P, _P]),
    "vs_shells_free": (c_int, [_P]),
    "vs_shells_num_layers": (c_int, [_P]),
    "vs_shells_info": (c_int, [_P, c_int, _P, _P]),
    "vs_shells_overflowed": (c_int, [_P]),
    "vs_shells_trace": (c_int, [_P, _P, _P, _I64, c_int, c_int, _P, _P, _P, _P, _P]),
    "vs_shells_expand": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "vs_shells_sample_normals": (c_int, [_P, _P, _P, _I64, _P, _P, _P]),
    "vs_shells_sample_uvs": (c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "vs_mlp_blob_bytes": (_I64, [c_int, _P]),
    "vs_mlp_pack": (c_int, [c_int, _P, _P, _P, _P, _P]),
    "vs_mlp_forward": (c_int, [c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _I64, _P, c_int, _P]),
    "vs_mlp_stash_bytes": (_I64, [c_int, _P, _I64]),
    "vs_importance_sample": (c_int, [_P, _P, _P, _P, _P, _I64, _I64, c_int, ctypes.c_uint64, ctypes.c_uint64, c_int, _P, _P, _P, _P, _P]),
    "vs_combine_offsets": (c_int, [_P, _P, _I64, _P, _P, _P, _P]),
    "vs_combine_merge": (c_int, [_I64, c_float, c_int] + [_P] * 19 + [_P]),
    "vs_mlp_num_params": (_I64, [c_int, _P]),
    "vs_mlp_backward_workspace_bytes": (_I64, [c_int, _P, c_int, c_int, c_int, _I64]),
    "vs_mlp_backward_stashed": (c_int, [c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _I64, _P, _P]),
    "vs_permuto_output_dims": (c_int, [c_int, c_int, c_int]),
    "vs_permuto_forward": (c_int, [c_int, c_int, _I64, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _P, _I64, _P, _P]),
    "vs_permuto_backward": (c_int, [c_int, c_int, _I64, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _P, _P, _I64, _P, _P]),
    "vs_permuto_backward_keyed": (c_int, [c_int, c_int, _I64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _I64, _P, _P, _I64, _P, _P]),
    "vs_hashgrid_levels": (_I64, [c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    "vs_hashgrid_forward": (c_int, [c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, _P, _P, _P, _I64, _P, _P]),
    "vs_hashgrid_backward": (c_int, [c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, _P, _P, _P, _I64, _P, _P]),
    "vs_mlp_forward_raw": (c_int, [c_int, _P, _P, c_int, _P, _P, _P, _I64, _P, _P]),
    "vs_mlp_backward_stashed_raw": (c_int, [c_int, _P, _P, _P, c_int, _P, _P, _P, c_int, _P, _I64, _P, _P]),
    "vs_shtex_combine_forward": (c_int, [c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "vs_shtex_combine_backward": (c_int, [c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "vs_segment_offsets": (c_int, [_P, _I64, _P, _P, _P, _P]),
    "vs_sampler_fg_count": (c_int, [_P, _P, _P, _P, c_float, c_int, c_int, ctypes.c_uint64, ctypes.c_uint64, c_int, c_int, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_sampler_fg_write": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_sampler_bg": (c_int, [_P, _P, _P, c_float, c_int, ctypes.c_uint64, ctypes.c_uint64, c_int, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_sampler_contract": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _I64, _P]),
    "vs_occgrid_points": (c_int, [_P, c_int, _P, c_int, ctypes.c_uint64, ctypes.c_uint64, c_int, _P, _I64, _P]),
    "vs_occgrid_update_values": (c_int, [_P, _P, c_float, _P, _I64, _P]),
    "vs_occgrid_update_occupancy_density": (c_int, [_P, c_int, _P, c_float, c_int, _P, _P, _I64, _P]),
    "vs_occgrid_update_occupancy_sdf": (c_int, [_P, c_int, _P, _P, c_float, _P, _P, _I64, _P]),
    "vs_occgrid_first_sample_start": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_occgrid_advance_to_next_occupied": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_occgrid_rays_t_near_t_far": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_occgrid_check_occupancy": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _I64, _P]),
    "vs_mlp_backward": (c_int, [c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, _P, _I64, _P, c_int, _P]),
    "vs_baked_texture_shade": (c_int, [_P] * 7 + [c_int, c_int, c_int, _P] + [_P] * 6 + [_I64, _P]),
    "vs_blend_l1_loss": (c_int, [_P] * 9 + [_I64, _P]),
}

_lib = None


class VolsurfsB200Error(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the shared object once and attach the prototypes. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VolsurfsB200Error(
            f"{LIB_PATH} not found: build it with `python -m volsurfs_b200.build` "
            "(there is no CPU or PyTorch fallback for the volsurfs_b200 operators)"
        )
    handle = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as exc:  # pragma: no cover - build/ABI mismatch
            raise VolsurfsB200Error(f"{LIB_PATH} does not export {name}; rebuild the extension") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    """Raise on a non-zero return code of the C ABI."""
    if code != 0:
        msg = lib().vs_error_string(int(code))
        raise VolsurfsB200Error(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
