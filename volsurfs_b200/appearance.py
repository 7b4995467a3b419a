"""Fused appearance head: the B200 counterpart of the reference's legacy ``RGB`` model (volsurfs_py/models/rgb.py:13-168) with its
``MLP`` (models/mlp.py:8-69) and ``SHEncoder`` (encodings/sphericalharmonics.py:36-153), plus the alpha decay of
volsurfs_py/methods/volsurfs.py:583-594, evaluated for every packed layer hit in ONE tcgen05 kernel (csrc/mlp.cu).

The positional encoding (permutohedral hash) is the stage before this one and is an input here: ``pos_features`` [S, F]."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import check, ptr
from .volsurfs import _stream


class AppearanceHead(torch.nn.Module):
    """[pos_features | SH_deg(dirs) | normals?] -> Linear+act x len(hidden) -> Linear -> sigmoid (* alpha decay).

    Parameters are fp32 ``torch.nn.Linear`` modules initialised like the reference (PyTorch defaults, mlp.py:54-69); the
    kernel consumes an fp16 tensor-core packing that is refreshed whenever a parameter changes."""

    def __init__(self, pos_dim: int = 51, hidden=(128, 128, 64), out_dim: int = 3, sh_degree: int = 3, normal_dep: bool = False,
                 activation: str = "gelu", alpha_decay: bool = False):
        super().__init__()
        assert activation in ("gelu", "relu")
        self.pos_dim, self.out_dim, self.sh_degree = pos_dim, out_dim, sh_degree
        self.normal_dep, self.activation, self.alpha_decay = bool(normal_dep), activation, bool(alpha_decay)
        n_sh = 0 if sh_degree < 0 else (sh_degree + 1) ** 2
        self.dims = [pos_dim + n_sh + (3 if normal_dep else 0)] + list(hidden) + [out_dim]
        self.layers = torch.nn.ModuleList(torch.nn.Linear(self.dims[i], self.dims[i + 1]) for i in range(len(self.dims) - 1))
        self._blob = None
        self._blob_key = None

    def load_linear_stack(self, weights, biases):
        with torch.no_grad():
            for lin, W, b in zip(self.layers, weights, biases):
                lin.weight.copy_(W)
                lin.bias.copy_(b)
        return self

    def _dims_c(self):
        return (ctypes.c_int * len(self.dims))(*self.dims)

    def packed(self):
        """fp16 UMMA-layout weights + fp32 biases (device blob), re-packed when any parameter was modified"""
        L = _lib.lib()
        params = [p for lin in self.layers for p in (lin.weight, lin.bias)]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._blob is None or key != self._blob_key:
            n = len(self.layers)
            nbytes = int(L.vs_mlp_blob_bytes(n, self._dims_c()))
            if nbytes < 0:
                check(nbytes, "vs_mlp_blob_bytes")
            dev = self.layers[0].weight.device
            if self._blob is None or self._blob.numel() != nbytes or self._blob.device != dev:
                self._blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ws = [lin.weight.detach().float().contiguous() for lin in self.layers]
            bs = [lin.bias.detach().float().contiguous() for lin in self.layers]
            wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
            bp = (ctypes.c_void_p * n)(*[b.data_ptr() for b in bs])
            check(L.vs_mlp_pack(n, self._dims_c(), wp, bp, ptr(self._blob), _stream()), "vs_mlp_pack")
            self._blob_key = key
        return self._blob

    def forward(self, pos_features, dirs, normals=None, n_valid_dev=None, out=None, _variant: int = 0):
        """pos_features [S,F] f32, dirs [S,3] f32, normals [S,3] f32 (needed for normal_dep / alpha_decay) -> [S,out_dim] f32.

        Differentiable w.r.t. the Linear parameters and ``pos_features`` (fused backward kernel, csrc/mlp_bwd.cu); ``dirs`` and
        ``normals`` carry no gradient, as in the reference (rgb.py:123-124, volsurfs.py:583-594)."""
        if torch.is_grad_enabled() and (pos_features.requires_grad or any(p.requires_grad for p in self.parameters())):
            params = [p for lin in self.layers for p in (lin.weight, lin.bias)]
            return _HeadFunction.apply(self, pos_features, dirs, normals, n_valid_dev, _variant, *params)
        with torch.no_grad():
            return self._forward_impl(pos_features, dirs, normals, n_valid_dev, out, _variant)

    def stash_bytes(self, n_samples: int) -> int:
        n = int(_lib.lib().vs_mlp_stash_bytes(len(self.layers), self._dims_c(), int(n_samples)))
        if n < 0:
            check(n, "vs_mlp_stash_bytes")
        return n

    def new_stash(self, n_samples: int, device=None):
        """activation stash for a training-mode forward of ``n_samples`` samples (per 128-sample tile: every layer's fp16 operand
        and every hidden layer's activation derivative)"""
        dev = device if device is not None else self.layers[0].weight.device
        return torch.empty(max(self.stash_bytes(n_samples), 16), dtype=torch.uint8, device=dev)

    @torch.no_grad()
    def forward_train(self, pos_features, dirs, normals=None, n_valid_dev=None, out=None, stash=None):
        """forward that also fills ``stash`` (allocated when None) for ``backward_into(..., stash=, fwd_out=)``; returns (out, stash)"""
        if stash is None:
            stash = self.new_stash(int(pos_features.shape[0]), pos_features.device)
        out = self._forward_impl(pos_features, dirs, normals, n_valid_dev, out, 0, stash)
        return out, stash

    def _forward_impl(self, pos_features, dirs, normals=None, n_valid_dev=None, out=None, _variant: int = 0, stash=None):
        S = int(pos_features.shape[0])
        pos_features = pos_features.contiguous()
        assert pos_features.dtype == torch.float32 and pos_features.shape[1] == self.pos_dim
        if out is None:
            out = torch.empty((S, self.out_dim), dtype=torch.float32, device=pos_features.device)
        blob = self.packed()
        check(
            _lib.lib().vs_mlp_forward(
                len(self.layers), self._dims_c(), ptr(blob), self.pos_dim, self.sh_degree, int(self.normal_dep),
                1 if self.activation == "gelu" else 0, int(self.alpha_decay), ptr(pos_features),
                ptr(None if dirs is None else dirs.contiguous()), ptr(None if normals is None else normals.contiguous()), ptr(out),
                ptr(stash), S, ptr(n_valid_dev), int(_variant), _stream(),
            ),
            "vs_mlp_forward",
        )
        return out

    # ---- backward ----------------------------------------------------------------------------------------------------
    def num_params(self) -> int:
        return int(_lib.lib().vs_mlp_num_params(len(self.layers), self._dims_c()))

    def split_flat(self, flat):
        """views of a flat [W_0 | b_0 | W_1 | b_1 ...] vector shaped like the Linear parameters"""
        out, o = [], 0
        for lin in self.layers:
            n, k = lin.weight.shape
            out.append(flat[o:o + n * k].view(n, k))
            o += n * k
            out.append(flat[o:o + n])
            o += n
        return out

    @torch.no_grad()
    def backward_into(self, pos_features, dirs, normals, d_out, d_params, d_pos=None, accumulate=False, n_valid_dev=None,
                      _variant: int = 0, stash=None, fwd_out=None):
        """One launch sequence of the fused backward: fills ``d_params`` (flat fp32, ``num_params()`` entries) and, when given,
        ``d_pos`` [S, pos_dim].  With ``stash`` + ``fwd_out`` (from ``forward_train``) the kernel streams the saved activations
        (HBM-bound); without them it recomputes the forward pass from the inputs."""
        L = _lib.lib()
        S = int(pos_features.shape[0])
        pos_features = pos_features.contiguous()
        d_out = d_out.contiguous()
        assert d_out.dtype == torch.float32 and tuple(d_out.shape) == (S, self.out_dim)
        assert d_params.dtype == torch.float32 and d_params.numel() == self.num_params() and d_params.is_contiguous()
        sh = self.sh_degree
        ws_bytes = int(L.vs_mlp_backward_workspace_bytes(len(self.layers), self._dims_c(), self.pos_dim, sh, int(self.normal_dep), S))
        if ws_bytes < 0:
            check(ws_bytes, "vs_mlp_backward_workspace_bytes")
        ws = getattr(self, "_bwd_ws", None)
        if ws is None or ws.numel() < ws_bytes or ws.device != pos_features.device:
            ws = self._bwd_ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pos_features.device)
        if stash is not None:
            assert fwd_out is not None and fwd_out.is_contiguous() and tuple(fwd_out.shape) == (S, self.out_dim)
            check(
                L.vs_mlp_backward_stashed(
                    len(self.layers), self._dims_c(), ptr(self.packed()), ptr(stash), self.pos_dim, sh, int(self.normal_dep),
                    1 if self.activation == "gelu" else 0, int(self.alpha_decay), ptr(None if dirs is None else dirs.contiguous()),
                    ptr(None if normals is None else normals.contiguous()), ptr(fwd_out), ptr(d_out), ptr(d_pos), ptr(d_params),
                    int(bool(accumulate)), ptr(ws), S, ptr(n_valid_dev), _stream(),
                ),
                "vs_mlp_backward_stashed",
            )
            return d_params, d_pos
        check(
            L.vs_mlp_backward(
                len(self.layers), self._dims_c(), ptr(self.packed()), self.pos_dim, sh, int(self.normal_dep),
                1 if self.activation == "gelu" else 0, int(self.alpha_decay), ptr(pos_features),
                ptr(None if dirs is None else dirs.contiguous()), ptr(None if normals is None else normals.contiguous()), ptr(d_out),
                ptr(d_pos), ptr(d_params), int(bool(accumulate)), ptr(ws), S, ptr(n_valid_dev), int(_variant), _stream(),
            ),
            "vs_mlp_backward",
        )
        return d_params, d_pos


class _HeadFunction(torch.autograd.Function):
    """autograd glue: forward = vs_mlp_forward, backward = vs_mlp_backward (parameter gradients returned as views of one flat buffer)"""

    @staticmethod
    def forward(ctx, head, pos_features, dirs, normals, n_valid_dev, variant, *params):
        # variant bit 2 (debug / tests): recompute-in-backward path instead of the activation stash
        stash = None if (variant & 4) else head.new_stash(int(pos_features.shape[0]), pos_features.device)
        out = head._forward_impl(pos_features, dirs, normals, n_valid_dev, None, variant & 3, stash)
        ctx.stash = stash
        ctx.head = head
        ctx.variant = variant
        ctx.n_valid_dev = n_valid_dev
        ctx.need_pos = pos_features.requires_grad
        ctx.save_for_backward(pos_features, dirs, normals, out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        head = ctx.head
        pos_features, dirs, normals, fwd_out = ctx.saved_tensors
        flat = torch.empty(head.num_params(), dtype=torch.float32, device=g_out.device)
        d_pos = torch.zeros_like(pos_features) if ctx.need_pos and ctx.n_valid_dev is not None else (
            torch.empty_like(pos_features) if ctx.need_pos else None)
        head.backward_into(pos_features, dirs, normals, g_out, flat, d_pos, False, ctx.n_valid_dev, ctx.variant & 3, ctx.stash,
                           fwd_out if ctx.stash is not None else None)
        ctx.head = ctx.stash = None
        return (None, d_pos, None, None, None, None, *head.split_flat(flat))
