"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement (torch + numpy) of the reference's DEFAULT appearance (SURVEY.md 8a row a6' / 8f row 4): SH neural textures.

* SHNeuralTextures.forward ........ volsurfs_py/models/sh_neural_textures.py:64-97
* NeuralTexture.forward ........... volsurfs_py/models/neural_texture.py:81-197 (anchor / lerp modes, align_to_webgl, squeeze,
                                    8-bit quantisation with the straight-through estimator of volsurfs_py/utils/math.py:5-19,
                                    fp16 re-expansion to val_range, 4-corner lerp)
* uv helpers ...................... submodules/mvdatasets/mvdatasets/utils/images.py:30-117
* SHEncoder.eval .................. volsurfs_py/encodings/sphericalharmonics.py:156-229 (mixed fp16 / fp32 evaluation)
* texture network ................. tiny-cuda-nn (UN-VENDORED, no version pin: README.md:46; call site neural_texture.py:54-79):
                                    ``HashGrid`` encoding (2-D, 16 levels x 2 features, 2^15 entries, base 16, scale 1.5, linear
                                    interpolation) -> ``FullyFusedMLP`` (64 neurons, 2 hidden layers, ReLU, no biases, linear output).

PIN STATUS.  The glue (everything except the two tiny-cuda-nn modules) is pinned bit-exact by tests/golden/shtex_*.npz, produced
by importing the reference's own ``NeuralTexture`` / ``SHNeuralTextures`` classes with ``tinycudann`` replaced by a stub that calls
the restatement below (tests/golden/make_golden_shtex.py).  The tiny-cuda-nn arithmetic itself is **parity unpinned**: the package
is absent and needs a GPU.  The restatement follows its published algorithm (grid.h: grid_scale / grid_resolution / grid_index /
coherent-prime hash / kernel_grid; fully_fused_mlp.cu: fp16 weights and activations) and fixes the roundings it leaves to the
compiler as the contract of this repository:
    pos   = fl32(fl32(scale * x) + 0.5)         (no FMA contraction)
    w_c   = fl32 products of (1 - frac) / frac in dimension order
    f     = fp16( sum_c fl32(w_c * fp16(table[c])) )   fp32 accumulation in corner order 0..3 (tiny-cuda-nn accumulates in fp16)
    MLP   : fp16 operands, fp32 accumulation, activations rounded to fp16, output rounded to fp16
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .appearance import C0, C1, C2, C3

PRIME_Y = 2654435761  # tiny-cuda-nn coherent prime hash: primes[0] = 1, primes[1] = 2654435761


# ---- tiny-cuda-nn HashGrid geometry -------------------------------------------------------------------------------------------------
def hashgrid_levels(n_levels=16, log2_hashmap_size=15, base_resolution=16, per_level_scale=1.5, n_pos_dims=2):
    """grid.h: grid_scale = exp2(level * log2(per_level_scale)) * base - 1; grid_resolution = ceil(scale) + 1; the level owns
    min(next_multiple(resolution^D, 8), 2^log2_hashmap_size) entries.  The scale is evaluated in double and rounded to fp32 (the
    same expression the C ABI evaluates on the host, csrc/shtex.cu:vs_hashgrid_levels)."""
    levels, offset = [], 0
    for lvl in range(n_levels):
        scale = np.float32(math.pow(float(per_level_scale), float(lvl)) * float(base_resolution) - 1.0)
        res = int(math.ceil(float(scale))) + 1
        dense = res ** n_pos_dims
        size = min((min(dense, 2 ** 31 - 1) + 7) // 8 * 8, 1 << log2_hashmap_size)
        levels.append(dict(scale=scale, res=res, size=size, offset=offset, hashed=dense > size))
        offset += size
    return levels, offset


def hashgrid_corner_terms(x: np.ndarray, lvl: dict):
    """indices [R,4] (entry inside the level) and fp32 weights [R,4] of the 4 interpolation corners (kernel_grid, grid_index)"""
    x = np.asarray(x, np.float32)
    scale = np.float32(lvl["scale"])
    pos = (x * scale).astype(np.float32) + np.float32(0.5)
    fl = np.floor(pos)
    grid = fl.astype(np.int64).astype(np.uint32)  # (uint32_t)(int)floorf(pos)
    frac = (pos - fl).astype(np.float32)
    one = np.float32(1.0)
    idx = np.empty((x.shape[0], 4), np.int64)
    w = np.empty((x.shape[0], 4), np.float32)
    res, size = np.uint32(lvl["res"]), np.uint32(lvl["size"])
    for c in range(4):
        gx = grid[:, 0] + np.uint32(c & 1)
        gy = grid[:, 1] + np.uint32((c >> 1) & 1)
        wx = frac[:, 0] if (c & 1) else (one - frac[:, 0])
        wy = frac[:, 1] if (c & 2) else (one - frac[:, 1])
        w[:, c] = (wx * wy).astype(np.float32)
        if lvl["hashed"]:
            h = gx ^ (gy * np.uint32(PRIME_Y))  # uint32 wrap-around
        else:
            h = gx + gy * res
        idx[:, c] = (h % size).astype(np.int64)
    return idx, w


class RoundHalfSTE(torch.autograd.Function):
    """x -> fp16 -> fp32, identity gradient (a rounding inside a network whose arithmetic is fp16)"""

    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


def hashgrid_forward(table: torch.Tensor, x: torch.Tensor, levels) -> torch.Tensor:
    """table [n_entries, 2] fp32 master parameters, x [R,2] fp32 in [0,1] -> features [R, 2*n_levels] fp32 holding fp16 values
    (column = level*2 + feature).  Differentiable w.r.t. ``table`` (fp32 gradient of the fp16-rounded forward)."""
    xn = x.detach().cpu().numpy().astype(np.float32)
    cols = []
    for lvl in levels:
        idx, w = hashgrid_corner_terms(xn, lvl)
        idx_t = torch.from_numpy(idx + lvl["offset"])
        w_t = torch.from_numpy(w)
        acc = None
        for c in range(4):
            val = RoundHalfSTE.apply(table[idx_t[:, c]])  # [R,2]
            term = w_t[:, c:c + 1] * val
            acc = term if acc is None else acc + term
        cols.append(RoundHalfSTE.apply(acc))
    return torch.cat(cols, 1)


def mlp_half_forward(x: torch.Tensor, weights) -> torch.Tensor:
    """FullyFusedMLP restatement: fp16 weights/activations, fp32 accumulation, ReLU hidden layers, linear fp16 output, no biases.
    weights[l]: [out_l, in_l] fp32 master parameters.  Returns fp32 holding fp16 values."""
    h = x
    for i, W in enumerate(weights):
        h = h @ RoundHalfSTE.apply(W).t()
        if i < len(weights) - 1:
            h = torch.relu(h)
        h = RoundHalfSTE.apply(h)
    return h


class TextureNet:
    """the ``torch.nn.Sequential(tcnn.Encoding, tcnn.Network)`` of neural_texture.py:63-79 (restated)"""

    def __init__(self, n_out: int, seed: int = 0, n_levels=16, log2_hashmap_size=15, base_resolution=16, per_level_scale=1.5,
                 n_neurons=64, n_hidden_layers=2, table_init=1e-4):
        self.levels, self.n_entries = hashgrid_levels(n_levels, log2_hashmap_size, base_resolution, per_level_scale)
        self.seed, self.table_init = seed, table_init
        g = torch.Generator().manual_seed(seed)
        self.table = ((torch.rand(self.n_entries, 2, generator=g) * 2 - 1) * table_init).requires_grad_()  # tcnn: U(-1e-4, 1e-4)
        dims = [2 * n_levels] + [n_neurons] * n_hidden_layers + [n_out]
        self.weights = []
        for i in range(len(dims) - 1):
            bound = math.sqrt(6.0 / (dims[i] + dims[i + 1]))  # tcnn: xavier uniform
            self.weights.append(((torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * bound).requires_grad_())
        self.n_out = n_out

    def parameters(self):
        return [self.table] + self.weights

    def features(self, uv):
        return hashgrid_forward(self.table, uv, self.levels)

    def raw(self, uv):
        """fp32 tensor holding the fp16 network output (what ``self.model(uv).float()`` is in the reference)"""
        return mlp_half_forward(self.features(uv), self.weights)

    def __call__(self, uv):
        return HalfCast.apply(self.raw(uv))

    @torch.no_grad()
    def min_abs_preactivation(self, uv):
        """smallest |hidden pre-activation| over a batch: a ReLU input within rounding noise of 0 makes the activation DERIVATIVE depend
        on the summation order (fixtures assert there is none, like the "no exact-t ties" rule of the ray-trace fixtures)"""
        h, m = self.features(uv), float("inf")
        for W in self.weights[:-1]:
            x = h @ W.half().float().t()
            m = min(m, float(x.abs().min()))
            h = torch.relu(x).half().float()
        return m


class HalfCast(torch.autograd.Function):
    """fp32 values that are exactly representable in fp16 -> a torch.half tensor (tiny-cuda-nn returns fp16); gradient passes"""

    @staticmethod
    def forward(ctx, x):
        return x.half()

    @staticmethod
    def backward(ctx, g):
        return g.float()


# ---- mvdatasets uv helpers (images.py:30-117) -----------------------------------------------------------------------------------
def _corners(uv_pix_nn):
    offs = [torch.tensor([0, 0]), torch.tensor([1, 0]), torch.tensor([0, 1]), torch.tensor([1, 1])]
    return torch.stack([uv_pix_nn + o for o in offs], dim=1)  # images.py:30-43


class RoundSTE(torch.autograd.Function):
    """volsurfs_py/utils/math.py:5-19"""

    @staticmethod
    def forward(ctx, x):
        return x.round()

    @staticmethod
    def backward(ctx, g):
        return g


def texel_queries(uv_coords, res, anchor=False, lerp=True, align_to_webgl=True):
    """neural_texture.py:96-150: normalised uv of every network query ([S,2] anchor / [4S,2] lerp) and the lerp weights [S,4,1] (or None)"""
    res_t = torch.tensor(res).long()
    flip = torch.flip(res_t, dims=[0])  # (width, height)
    if anchor:
        uv_pix = (uv_coords * flip).floor().long()  # images.py:64-67
        if align_to_webgl:
            width = res_t[1].item()
            temp = uv_pix[:, 0].clone()
            uv_pix[:, 0] = (width - 1) - uv_pix[:, 1]
            uv_pix[:, 1] = temp
        return (uv_pix.float() + 0.5) / flip, None  # images.py:84-94
    if not lerp:
        raise ValueError("NeuralTexture should be either anchor or lerp")
    uv_nn = uv_coords * flip  # images.py:54-61
    if align_to_webgl:
        width = res_t[1].item()
        temp = uv_nn[:, 0].clone()
        uv_nn[:, 0] = width - uv_nn[:, 1]
        uv_nn[:, 1] = temp
    corners = _corners((uv_nn - 0.5).floor()) + 0.5  # images.py:70-81  [S,4,2]
    diff = uv_nn - corners[:, 0, :]  # images.py:97-117
    lerp_weights = torch.zeros((uv_nn.shape[0], 4))
    lerp_weights[:, 0] = (1.0 - diff[:, 0]) * (1.0 - diff[:, 1])
    lerp_weights[:, 1] = diff[:, 0] * (1.0 - diff[:, 1])
    lerp_weights[:, 2] = (1.0 - diff[:, 0]) * diff[:, 1]
    lerp_weights[:, 3] = diff[:, 0] * diff[:, 1]
    return corners.reshape(-1, 2) / flip, lerp_weights.unsqueeze(-1)  # images.py:46-51


def texture_from_network_output(output, n_out, val_range, lerp_weights, quantize_output=True, squeeze_output=True):
    """neural_texture.py:152-195: fp16 network output [rows, n_out] -> fp32 texture values [S, n_out]"""
    output = output.float()
    if squeeze_output:
        output = torch.sigmoid(output)
        if quantize_output:
            output = output * 255.0
            output = RoundSTE.apply(output)
            output = output / 255.0
    output = output.half()
    if squeeze_output:
        output = val_range[0] + (val_range[1] - val_range[0]) * output
    if lerp_weights is not None:
        output = output.reshape(-1, 4, n_out)
        output = (output * lerp_weights).sum(dim=1)
    return output.float()


def neural_texture_forward(net, uv_coords, res, val_range, anchor=False, lerp=True, quantize_output=True, squeeze_output=True,
                           align_to_webgl=True, keep=None):
    """neural_texture.py:81-197 (bake=False).  res = [height, width]; returns fp32 [S, n_out].  ``keep`` (a list) receives the fp16
    network output tensor (with retain_grad) for stage-wise checks."""
    uv_, lerp_weights = texel_queries(uv_coords, res, anchor, lerp, align_to_webgl)
    output = net(uv_)  # fp16
    if keep is not None:
        if output.requires_grad:
            output.retain_grad()
        keep.append(output)
    return texture_from_network_output(output, net.n_out, val_range, lerp_weights, quantize_output, squeeze_output)


def sh_eval(sh, dirs, degree):
    """sphericalharmonics.py:156-229 (degree <= 3 here).  sh [..., C, (degree+1)^2] (fp16 in the reference), dirs [..., 3] fp32."""
    result = C0 * sh[..., 0]
    if degree > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if degree > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6]
                      + C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if degree > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10]
                          + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12]
                          + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14]
                          + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


DEG_NR_COEFFS = [1, 3, 5, 7]


def sh_neural_textures_forward(nets, uv_coords, view_dirs, sh_deg=3, nr_channels=3, sh_range=(15.0, 15.0, 15.0, 15.0),
                               deg_res=(2048, 1024, 512, 256), anchor=False, lerp=True, quantize_output=True, squeeze_output=True,
                               align_to_webgl=True, keep=None):
    """sh_neural_textures.py:64-97.  nets[deg] is the degree's TextureNet (n_out = nr_channels * (2 deg + 1))."""
    n = uv_coords.shape[0]
    nr_coeffs = sum(DEG_NR_COEFFS[: sh_deg + 1])
    output = torch.zeros(n, nr_channels, nr_coeffs)
    written = 0
    for deg in range(sh_deg + 1):
        res = neural_texture_forward(nets[deg], uv_coords, [deg_res[deg], deg_res[deg]], (-sh_range[deg], sh_range[deg]), anchor, lerp,
                                     quantize_output, squeeze_output, align_to_webgl, keep).reshape(n, nr_channels, -1)
        output[:, :, written:written + DEG_NR_COEFFS[deg]] = res
        written += DEG_NR_COEFFS[deg]
    if view_dirs is None:
        return output
    sh_coeffs = output.half()
    raw = sh_eval(sh_coeffs, view_dirs, sh_deg)
    return torch.sigmoid(raw).float()
