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

    @torch.no_grad()
    def forward(self, pos_features, dirs, normals=None, n_valid_dev=None, out=None, _variant: int = 0):
        """pos_features [S,F] f32, dirs [S,3] f32, normals [S,3] f32 (needed for normal_dep / alpha_decay) -> [S,out_dim] f32"""
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
                ptr(None if dirs is None else dirs.contiguous()), ptr(None if normals is None else normals.contiguous()), ptr(out), S,
                ptr(n_valid_dev), int(_variant), _stream(),
            ),
            "vs_mlp_forward",
        )
        return out
