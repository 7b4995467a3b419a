"""ORACLE (test infrastructure only — never imported by the product).

CPU restatement (torch, fp32 or fp64) of the legacy appearance head evaluated at every layer hit:

* SH direction encoding ..... volsurfs_py/encodings/sphericalharmonics.py:84-153 (hard-coded real SH polynomials up to degree 4,
                              directions NOT normalised, evaluated under no_grad: rgb.py:123-124)
* MLP ....................... volsurfs_py/models/mlp.py:8-52 (Linear + exact-erf GELU per hidden layer, linear last layer)
* RGB head .................. volsurfs_py/models/rgb.py:104-149: x = [pos_features | SH(dirs) | normals?] -> MLP -> sigmoid
* alpha decay ............... volsurfs_py/methods/volsurfs.py:583-594: alpha *= 2*sigmoid(10*clamp(-d.n, 0, 1)) - 1 (no_grad)

The positional encoding (permutohedral hash, SURVEY.md section 8f "next" row 1) is an INPUT here: ``pos_features`` [S,F].

Pinned by tests/golden/appearance_*.npz, produced by importing the reference's own ``MLP`` and ``SHEncoder`` classes and exec'ing the
alpha-decay lines (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
      -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892, 0.10578554691520431, -0.6690465435572892,
      0.47308734787878004, -1.7701307697799304, 0.6258357354491761]


def sh_encode(dirs: torch.Tensor, degree: int = 3) -> torch.Tensor:
    """sphericalharmonics.py:84-153"""
    assert 0 <= degree <= 4
    x, y, z = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    out = torch.zeros((dirs.shape[0], (degree + 1) ** 2), dtype=dirs.dtype, device=dirs.device)
    out[:, 0] = C0
    if degree > 0:
        out[:, 1] = -C1 * y
        out[:, 2] = C1 * z
        out[:, 3] = -C1 * x
    if degree > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        out[:, 4] = C2[0] * xy
        out[:, 5] = C2[1] * yz
        out[:, 6] = C2[2] * (2.0 * zz - xx - yy)
        out[:, 7] = C2[3] * xz
        out[:, 8] = C2[4] * (xx - yy)
    if degree > 2:
        out[:, 9] = C3[0] * y * (3 * xx - yy)
        out[:, 10] = C3[1] * xy * z
        out[:, 11] = C3[2] * y * (4 * zz - xx - yy)
        out[:, 12] = C3[3] * z * (2 * zz - 3 * xx - 3 * yy)
        out[:, 13] = C3[4] * x * (4 * zz - xx - yy)
        out[:, 14] = C3[5] * z * (xx - yy)
        out[:, 15] = C3[6] * x * (xx - 3 * yy)
    if degree > 3:
        out[:, 16] = C4[0] * xy * (xx - yy)
        out[:, 17] = C4[1] * yz * (3 * xx - yy)
        out[:, 18] = C4[2] * xy * (7 * zz - 1)
        out[:, 19] = C4[3] * yz * (7 * zz - 3)
        out[:, 20] = C4[4] * (zz * (35 * zz - 30) + 3)
        out[:, 21] = C4[5] * xz * (7 * zz - 3)
        out[:, 22] = C4[6] * (xx - yy) * (7 * zz - 1)
        out[:, 23] = C4[7] * xz * (xx - 3 * yy)
        out[:, 24] = C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))
    return out


def mlp_forward(x: torch.Tensor, weights, biases, activation: str = "gelu") -> torch.Tensor:
    """mlp.py:8-52 with last_layer_linear=True: weights[i] is [out_i, in_i] (torch.nn.Linear layout)"""
    act = torch.nn.functional.gelu if activation == "gelu" else torch.relu  # torch.nn.GELU() = exact erf (mlp.py:36)
    h = x
    for i, (W, b) in enumerate(zip(weights, biases)):
        h = torch.nn.functional.linear(h, W, b)
        if i < len(weights) - 1:
            h = act(h)
    return h


def head_forward(pos_features, dirs, normals, weights, biases, sh_degree=3, normal_dep=False, activation="gelu"):
    """rgb.py:104-149: concat [pos | SH(dirs) | normals] -> MLP -> sigmoid"""
    with torch.no_grad():
        enc = sh_encode(dirs, sh_degree)
    parts = [pos_features, enc.to(pos_features.dtype)]
    if normal_dep:
        parts.append(normals)
    return torch.sigmoid(mlp_forward(torch.cat(parts, 1), weights, biases, activation))


def alpha_decay(alpha, dirs, normals, threshold: float = 10.0):
    """volsurfs.py:583-594"""
    with torch.no_grad():
        dot = torch.sum(-dirs * normals, dim=1, keepdim=True).clamp(0.0, 1.0)
        decay = torch.sigmoid(threshold * dot) * 2.0 - 1.0
    return alpha * decay


def init_linear_stack(in_dim, hidden, out_dim, seed, dtype=torch.float32):
    """torch.nn.Linear default init (mlp.py:54-69 re-inits every layer with PyTorch defaults), seeded"""
    g = torch.Generator().manual_seed(seed)
    dims = [in_dim] + list(hidden) + [out_dim]
    Ws, bs = [], []
    for i in range(len(dims) - 1):
        bound = 1.0 / (dims[i] ** 0.5)
        Ws.append(((torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * bound).to(dtype))
        bs.append(((torch.rand(dims[i + 1], generator=g) * 2 - 1) * bound).to(dtype))
    return Ws, bs
