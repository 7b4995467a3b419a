"""SH neural textures — the reference's DEFAULT appearance (config/volsurfs/base_5.cfg:11-20; SURVEY.md 8a row a6').

Mirrors ``volsurfs_py.models.neural_texture.NeuralTexture`` (neural_texture.py:17-197) and
``volsurfs_py.models.sh_neural_textures.SHNeuralTextures`` (sh_neural_textures.py:8-97): same constructor arguments, same
``forward(uv_coords, view_dirs=None, iter_nr=None)`` and the same tensors out.  The two tiny-cuda-nn modules every texture owns
(HashGrid encoding + FullyFusedMLP, neural_texture.py:54-79) are replaced by ``TextureNetwork``: a hash-grid kernel feeding the
tcgen05 MLP kernel (csrc/shtex.cu, csrc/mlp*.cu); everything after the networks is one fused kernel (``vs_shtex_combine_*``).

No CPU fallback: every stage calls the C ABI of libvolsurfs_b200.so."""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from ._lib import check, ptr
from .volsurfs import _stream

DEG_NR_COEFFS = [1, 3, 5, 7]


class TextureNetwork(torch.nn.Module):
    """``torch.nn.Sequential(tcnn.Encoding(HashGrid), tcnn.Network(FullyFusedMLP))`` of neural_texture.py:54-79.

    Parameters: ``table`` [entries, 2] fp32 (tiny-cuda-nn initialises U(-1e-4, 1e-4)) and bias-free Linear weights
    [32->64->64->n_out] (xavier uniform); the kernels consume fp16 roundings of both, like tiny-cuda-nn."""

    def __init__(self, n_out: int, n_levels: int = 16, n_features_per_level: int = 2, log2_hashmap_size: int = 15, base_resolution: int = 16,
                 per_level_scale: float = 1.5, n_neurons: int = 64, n_hidden_layers: int = 2):
        super().__init__()
        assert n_features_per_level == 2, "the hash-grid kernel is specialised for 2 features per level"
        self.grid_cfg = (int(n_levels), int(log2_hashmap_size), int(base_resolution), float(per_level_scale))
        n_entries = int(_lib.lib().vs_hashgrid_levels(*self.grid_cfg, None, None, None, None))
        if n_entries < 0:
            check(n_entries, "vs_hashgrid_levels")
        self.n_entries = n_entries
        self.n_out = int(n_out)
        self.dims = [2 * n_levels] + [n_neurons] * n_hidden_layers + [self.n_out]
        self.table = torch.nn.Parameter((torch.rand(n_entries, 2) * 2 - 1) * 1e-4)
        ws = []
        for i in range(len(self.dims) - 1):
            bound = math.sqrt(6.0 / (self.dims[i] + self.dims[i + 1]))
            ws.append(torch.nn.Parameter((torch.rand(self.dims[i + 1], self.dims[i]) * 2 - 1) * bound))
        self.weights = torch.nn.ParameterList(ws)
        self._blob = None
        self._blob_key = None

    def level_table(self):
        """per-level (scale, resolution, entries, offset) as the kernels use them"""
        L = self.grid_cfg[0]
        scale = (ctypes.c_float * L)()
        res, size, off = (ctypes.c_int32 * L)(), (ctypes.c_int32 * L)(), (ctypes.c_int32 * L)()
        check(min(int(_lib.lib().vs_hashgrid_levels(*self.grid_cfg, scale, res, size, off)), 0), "vs_hashgrid_levels")
        return list(scale), list(res), list(size), list(off)

    def _dims_c(self):
        return (ctypes.c_int * len(self.dims))(*self.dims)

    def packed(self):
        L = _lib.lib()
        key = tuple((w.data_ptr(), w._version) for w in self.weights)
        if self._blob is None or key != self._blob_key:
            n = len(self.weights)
            nbytes = int(L.vs_mlp_blob_bytes(n, self._dims_c()))
            if nbytes < 0:
                check(nbytes, "vs_mlp_blob_bytes")
            dev = self.weights[0].device
            if self._blob is None or self._blob.numel() != nbytes or self._blob.device != dev:
                self._blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ws = [w.detach().float().contiguous() for w in self.weights]
            wp = (ctypes.c_void_p * n)(*[w.data_ptr() for w in ws])
            bp = (ctypes.c_void_p * n)(*[None] * n)  # tiny-cuda-nn networks have no biases
            check(L.vs_mlp_pack(n, self._dims_c(), wp, bp, ptr(self._blob), _stream()), "vs_mlp_pack")
            self._blob_key = key
        return self._blob

    # ---- stages (no autograd) ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, uv, mode: int, align: bool, res_hw, n_valid_dev=None):
        """features [rows, 2*n_levels] of every texel query (rows = S * 4 in lerp mode)"""
        S = int(uv.shape[0])
        rows = S * (4 if mode == 1 else 1)
        feat = torch.empty((rows, self.dims[0]), dtype=torch.float32, device=uv.device)
        check(_lib.lib().vs_hashgrid_forward(*self.grid_cfg, mode, int(align), int(res_hw[0]), int(res_hw[1]), ptr(uv), ptr(self.table.detach()),
                                             ptr(feat), S, ptr(n_valid_dev), _stream()), "vs_hashgrid_forward")
        return feat

    @torch.no_grad()
    def mlp_raw(self, feat, stash=None, n_valid_rows_dev=None):
        rows = int(feat.shape[0])
        out = torch.empty((rows, self.n_out), dtype=torch.float32, device=feat.device)
        check(_lib.lib().vs_mlp_forward_raw(len(self.weights), self._dims_c(), ptr(self.packed()), 0, ptr(feat), ptr(out), ptr(stash), rows,
                                            ptr(n_valid_rows_dev), _stream()), "vs_mlp_forward_raw")
        return out

    def new_stash(self, rows: int, device):
        n = int(_lib.lib().vs_mlp_stash_bytes(len(self.weights), self._dims_c(), int(rows)))
        if n < 0:
            check(n, "vs_mlp_stash_bytes")
        return torch.empty(max(n, 16), dtype=torch.uint8, device=device)

    @torch.no_grad()
    def backward_into(self, uv, mode, align, res_hw, stash, d_raw):
        """(d_table, [d_W...]) from the gradient of the raw network output"""
        L = _lib.lib()
        rows = int(d_raw.shape[0])
        S = int(uv.shape[0])
        n = len(self.weights)
        n_params = int(L.vs_mlp_num_params(n, self._dims_c()))
        ws_bytes = int(L.vs_mlp_backward_workspace_bytes(n, self._dims_c(), self.dims[0], -1, 0, rows))
        if ws_bytes < 0:
            check(ws_bytes, "vs_mlp_backward_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=uv.device)
        flat = torch.empty(n_params, dtype=torch.float32, device=uv.device)
        d_feat = torch.empty((rows, self.dims[0]), dtype=torch.float32, device=uv.device)
        check(L.vs_mlp_backward_stashed_raw(n, self._dims_c(), ptr(self.packed()), ptr(stash), 0, ptr(d_raw), ptr(d_feat), ptr(flat), 0, ptr(ws),
                                            rows, None, _stream()), "vs_mlp_backward_stashed_raw")
        d_table = torch.zeros_like(self.table)
        check(L.vs_hashgrid_backward(*self.grid_cfg, mode, int(align), int(res_hw[0]), int(res_hw[1]), ptr(uv), ptr(d_feat), ptr(d_table), S,
                                     None, _stream()), "vs_hashgrid_backward")
        d_ws, o = [], 0
        for w in self.weights:
            k = w.numel()
            d_ws.append(flat[o:o + k].view_as(w))
            o += k + w.shape[0]  # skip the (unused) bias gradient slot
        return d_table, d_ws


def _combine_args(sh_deg, nr_channels, mode, align, res_list, ranges, squeeze, quantize):
    res_hw = (ctypes.c_int * (2 * (sh_deg + 1)))(*[int(v) for r in res_list for v in r])
    lo = (ctypes.c_float * (sh_deg + 1))(*[float(r[0]) for r in ranges])
    hi = (ctypes.c_float * (sh_deg + 1))(*[float(r[1]) for r in ranges])
    return [sh_deg, nr_channels, mode, int(align), res_hw, lo, hi, int(squeeze), int(quantize)]


class _ShTexFunction(torch.autograd.Function):
    """networks of all degrees -> coefficient assembly -> (SH evaluation).  ``nets``/``geoms`` describe one texture per degree."""

    @staticmethod
    def forward(ctx, spec, uv, dirs, *params):
        nets, mode, align, res_list, ranges, squeeze, quantize, C, sh_deg, want_coeffs = spec
        L = _lib.lib()
        S = int(uv.shape[0])
        uv = uv.detach().float().contiguous()
        dirs_c = None if dirs is None else dirs.detach().float().contiguous()
        train = any(ctx.needs_input_grad[3:])  # False under torch.no_grad(): then no activation stash is written
        raws, stashes = [], []
        for g, net in enumerate(nets):
            feat = net.encode(uv, mode, align, res_list[g])
            stash = net.new_stash(int(feat.shape[0]), uv.device) if train else None
            raws.append(net.mlp_raw(feat, stash))
            stashes.append(stash)
        n_coeffs = sum(DEG_NR_COEFFS[: sh_deg + 1])
        coeffs = torch.empty((S, C, n_coeffs), dtype=torch.float32, device=uv.device) if want_coeffs else None
        out = None if want_coeffs else torch.empty((S, C), dtype=torch.float32, device=uv.device)
        rp = (ctypes.c_void_p * (sh_deg + 1))(*[r.data_ptr() for r in raws])
        check(L.vs_shtex_combine_forward(*_combine_args(sh_deg, C, mode, align, res_list, ranges, squeeze, quantize), ptr(uv),
                                         None if want_coeffs else ptr(dirs_c), rp, ptr(coeffs), ptr(out), S, None, _stream()),
              "vs_shtex_combine_forward")
        ctx.spec = spec
        ctx.stashes = stashes
        ctx.train = train
        ctx.save_for_backward(uv, dirs_c if dirs_c is not None else uv, *raws, *([] if want_coeffs else [out]))
        return coeffs if want_coeffs else out

    @staticmethod
    def backward(ctx, g):
        nets, mode, align, res_list, ranges, squeeze, quantize, C, sh_deg, want_coeffs = ctx.spec
        assert ctx.train, "backward through SHNeuralTextures needs a forward pass run with gradients enabled"
        L = _lib.lib()
        saved = ctx.saved_tensors
        uv, dirs_c = saved[0], saved[1]
        raws = list(saved[2:2 + sh_deg + 1])
        out = None if want_coeffs else saved[2 + sh_deg + 1]
        S = int(uv.shape[0])
        g = g.detach().float().contiguous()
        d_raws = [torch.empty_like(r) for r in raws]
        rp = (ctypes.c_void_p * (sh_deg + 1))(*[r.data_ptr() for r in raws])
        dp = (ctypes.c_void_p * (sh_deg + 1))(*[r.data_ptr() for r in d_raws])
        check(L.vs_shtex_combine_backward(*_combine_args(sh_deg, C, mode, align, res_list, ranges, squeeze, quantize), ptr(uv),
                                          None if want_coeffs else ptr(dirs_c), rp, ptr(out), None if want_coeffs else ptr(g),
                                          ptr(g) if want_coeffs else None, dp, S, None, _stream()), "vs_shtex_combine_backward")
        grads = []
        for gi, net in enumerate(nets):
            d_table, d_ws = net.backward_into(uv, mode, align, res_list[gi], ctx.stashes[gi], d_raws[gi])
            grads += [d_table, *d_ws]
        ctx.stashes = None
        return (None, None, None, *grads)


def _net_params(nets):
    return [p for net in nets for p in (net.table, *net.weights)]


class NeuralTexture(torch.nn.Module):
    """neural_texture.py:17-197 — one texture: uv -> [S, nr_channels] (fp32)."""

    def __init__(self, res, nr_channels, val_range=(0.0, 1.0), anchor=False, lerp=False, quantize_output=False, squeeze_output=False,
                 align_to_webgl=False):
        super().__init__()
        if isinstance(res, torch.Tensor):
            res = res.tolist()
        if not isinstance(res, (list, tuple)):
            raise ValueError("NeuralTexture res should be a list or a torch.Tensor.")  # reference: print + exit(1)
        if anchor and lerp:
            raise ValueError("NeuralTexture cannot anchor and lerp at the same time.")
        self.res = torch.tensor([int(res[0]), int(res[1])]).long()  # (height, width)
        self.nr_channels = int(nr_channels)
        self.anchor, self.lerp = bool(anchor), bool(lerp)
        self.quantize_output, self.squeeze_output = bool(quantize_output), bool(squeeze_output)
        self.val_range = (float(val_range[0]), float(val_range[1]))
        self.align_to_webgl = bool(align_to_webgl)
        self.model = TextureNetwork(self.nr_channels)

    @property
    def mode(self):
        if self.anchor:
            return 0
        if self.lerp:
            return 1
        raise ValueError("NeuralTexture should be either anchor or lerp or bake.")

    def forward(self, uv_coords, bake=False, iter_nr=None):
        res_hw = [int(self.res[0]), int(self.res[1])]
        if bake:
            # texel-centre queries returned before the fp16 expansion (neural_texture.py:83-86, 171-174): off the per-ray path, used
            # once per texture when baking; network kernels + three torch elementwise ops
            with torch.no_grad():
                uv = uv_coords.detach().float().contiguous()
                raw = self.model.mlp_raw(self.model.encode(uv, 2, False, res_hw))
                out = raw.half().float()
                if self.squeeze_output:
                    out = torch.sigmoid(out)
                    if self.quantize_output:
                        out = torch.round(out * 255.0) / 255.0
                return out
        spec = ([self.model], self.mode, self.align_to_webgl, [res_hw], [self.val_range], self.squeeze_output, self.quantize_output,
                self.nr_channels, 0, True)
        return _ShTexFunction.apply(spec, uv_coords, None, *_net_params([self.model])).reshape(-1, self.nr_channels)


class SHNeuralTextures(torch.nn.Module):
    """sh_neural_textures.py:8-97 — one NeuralTexture per SH degree; ``forward`` returns sigmoid(SH(view_dirs)) [S, nr_channels], or the
    coefficient tensor [S, nr_channels, (sh_deg+1)^2] when ``view_dirs`` is None."""

    def __init__(self, sh_deg=0, nr_channels=3, sh_range=[1.0, 5.0, 10.0, 20.0], anchor=False, lerp=False, deg_res=[2048, 1024, 512, 256],
                 quantize_output=False, squeeze_output=False, align_to_webgl=False):
        super().__init__()
        if sh_deg >= 4:
            raise ValueError("SHNeuralTextures only supports SH degrees up to 3.")  # reference: print + exit(1)
        if quantize_output and not squeeze_output:
            raise ValueError("quantize_output requires squeeze_output.")
        self.sh_deg, self.nr_channels = int(sh_deg), int(nr_channels)
        self.deg_res, self.sh_range = list(deg_res), list(sh_range)
        self.deg_nr_coeffs = list(DEG_NR_COEFFS)
        self.nr_coeffs = sum(self.deg_nr_coeffs[: self.sh_deg + 1])
        self.neural_textures = torch.nn.ModuleList(
            NeuralTexture(res=[self.deg_res[d], self.deg_res[d]], nr_channels=self.nr_channels * self.deg_nr_coeffs[d],
                          val_range=(-self.sh_range[d], self.sh_range[d]), anchor=anchor, lerp=lerp, quantize_output=quantize_output,
                          squeeze_output=squeeze_output, align_to_webgl=align_to_webgl)
            for d in range(self.sh_deg + 1))

    def forward(self, uv_coords, view_dirs=None, iter_nr=None):
        nts = list(self.neural_textures)
        nets = [nt.model for nt in nts]
        spec = (nets, nts[0].mode, nts[0].align_to_webgl, [[int(nt.res[0]), int(nt.res[1])] for nt in nts], [nt.val_range for nt in nts],
                nts[0].squeeze_output, nts[0].quantize_output, self.nr_channels, self.sh_deg, view_dirs is None)
        return _ShTexFunction.apply(spec, uv_coords, view_dirs, *_net_params(nets))
