"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCE LINES (run once, in the build container where
/root/reference is mounted; the resulting .npz files are committed, the reference sources are not).

    python tests/golden/make_golden.py

dense_composite_*.npz
    The dense K-layer compositing block of the volsurfs method is read from
    /root/reference/volsurfs_py/methods/volsurfs.py (lines 600-643: flip, [.half()], cumprod, shift, weights, sum;
    line 708: final composite), dedented and exec'd on seeded CPU tensors.  The fp32 variant drops the two
    ``.half()`` casts (lines 606-607, 705); the fp16 variant keeps them (reference-faithful).  Gradients come from
    torch autograd through those same lines.
"""
from __future__ import annotations

import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))

from volsurfs_b200.synthetic import dense_layers  # noqa: E402  (input generator only)


def reference_block(first: int, last: int, drop_half: bool) -> str:
    lines = (REF / "volsurfs_py/methods/volsurfs.py").read_text().splitlines()[first - 1:last]
    if drop_half:
        lines = [ln for ln in lines if ".half()" not in ln]
    return textwrap.dedent("\n".join(lines))


def run_reference_dense(alpha, rgb, rgb_bg, half: bool):
    """alpha [N,K,1], rgb [N,K,3] in mesh order -> dict of reference outputs (+ autograd grads)."""
    K = alpha.shape[1]
    a = alpha.clone().requires_grad_(True)
    c = rgb.clone().requires_grad_(True)
    ns = {
        "torch": torch,
        "self": types.SimpleNamespace(nr_meshes=K, profiler=None),
        "surfs_alpha": a,
        "surfs_rgb": c,
        "debug_ray_idx": None,
    }
    exec(reference_block(600, 643, drop_half=not half), ns)  # volsurfs.py:600-643
    ns["rgb_bg"] = rgb_bg.half() if half else rgb_bg         # volsurfs.py:705
    exec(reference_block(708, 708, drop_half=False), ns)     # volsurfs.py:708
    out = {
        "rgb_fg": ns["rgb_fg"], "bg_transmittance": ns["bg_transmittance"], "pred_rgb": ns["pred_rgb"],
        "weights": ns["surfs_blending_weights"], "transmittance": ns["surfs_transmittance"],
    }
    return out, a, c


def make_dense(name: str, n_rays: int, K: int, seed_offset: int):
    d = dense_layers(n_rays, K, seed_offset=seed_offset)
    rgb_bg = torch.ones(n_rays, 3)
    save = {"hit": d["hit"].numpy(), "alpha": d["alpha"].numpy(), "rgb": d["rgb"].numpy(), "z": d["z"].numpy(),
            "g_rgb": d["g_rgb"].numpy(), "g_bgT": d["g_bgT"].numpy(), "rgb_bg": rgb_bg.numpy()}
    for half in (False, True):
        out, a, c = run_reference_dense(d["alpha"], d["rgb"], rgb_bg, half)
        tag = "fp16" if half else "fp32"
        loss = (out["rgb_fg"].float() * d["g_rgb"]).sum() + (out["bg_transmittance"].float() * d["g_bgT"]).sum()
        loss.backward()
        for k, v in out.items():
            save[f"{tag}_{k}"] = v.detach().float().numpy()
        save[f"{tag}_d_alpha"] = a.grad.numpy()
        save[f"{tag}_d_rgb"] = c.grad.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **save)
    print("wrote", name, {k: v.shape for k, v in save.items() if k.startswith("fp32")})


# ---------------------------------------------------------------------------------------------------------------------
# appearance head: the reference's own MLP and SHEncoder classes (imported), alpha-decay lines (exec'd)
# ---------------------------------------------------------------------------------------------------------------------
def make_appearance(name: str, n: int, hidden, out_dim: int, seed: int, normal_dep: bool):
    sys.path.insert(0, str(REF))
    sys.modules.setdefault("permutohedral_encoding", types.ModuleType("permutohedral_encoding"))  # imported, never used by SHEncoder
    import scipy.special

    if not hasattr(scipy.special, "sph_harm"):  # the reference imports a name newer SciPy removed; SHEncoder.__call__ never uses it
        scipy.special.sph_harm = scipy.special.sph_harm_y
    from volsurfs_py.encodings.sphericalharmonics import SHEncoder  # reference class
    from volsurfs_py.models.mlp import MLP                            # reference class

    g = torch.Generator().manual_seed(seed)
    F = 51                                                  # permutohash output dim (encodings/permutohash.py:38-41)
    pos = torch.rand(n, F, generator=g) * 2 - 1
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    normals = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    g_out = torch.randn(n, out_dim, generator=g)
    enc = SHEncoder(input_dim=3, degree=3)
    in_dim = F + enc.output_dim + (3 if normal_dep else 0)
    torch.manual_seed(seed)
    mlp = MLP(in_dim, list(hidden) + [out_dim], last_layer_linear=True)   # rgb.py:93-97
    pos_g = pos.clone().requires_grad_(True)
    with torch.set_grad_enabled(False):                                   # rgb.py:123-124
        sh = enc(dirs, iter_nr=None)
    data = torch.cat([pos_g, sh] + ([normals] if normal_dep else []), 1)  # rgb.py:118-131
    pre = mlp(data)
    out = torch.sigmoid(pre)                                              # rgb.py:147
    # alpha decay: exec volsurfs.py:582-595 on the reference's variable names
    lines = (REF / "volsurfs_py/methods/volsurfs.py").read_text().splitlines()[581:595]
    ns = {"torch": torch, "self": types.SimpleNamespace(with_alpha_decay=True), "surfs_points_alpha_pred": out[:, :1],
          "rays_d": dirs, "hits": slice(None), "surfs_normals": normals.unsqueeze(1), "i": 0}
    exec(textwrap.dedent("\n".join(lines)), ns)
    decayed = ns["surfs_points_alpha_pred"]
    (out * g_out).sum().backward()
    linears = [m for m in mlp.layers if isinstance(m, torch.nn.Linear)]
    save = {"pos": pos.numpy(), "dirs": dirs.numpy(), "normals": normals.numpy(), "g_out": g_out.numpy(), "sh": sh.numpy(),
            "pre": pre.detach().numpy(), "out": out.detach().numpy(), "alpha_decayed": decayed.detach().numpy(),
            "d_pos": pos_g.grad.numpy(), "hidden": np.array(hidden), "normal_dep": np.array(normal_dep)}
    for i, lin in enumerate(linears):
        save[f"W{i}"] = lin.weight.detach().numpy()
        save[f"b{i}"] = lin.bias.detach().numpy()
        save[f"dW{i}"] = lin.weight.grad.numpy()
        save[f"db{i}"] = lin.bias.grad.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **save)
    print("wrote", name, in_dim, [tuple(l.weight.shape) for l in linears])


if __name__ == "__main__":
    assert REF.exists(), "run in the container that mounts /root/reference"
    make_dense("dense_composite_k5", 96, 5, seed_offset=101)
    make_dense("dense_composite_k1", 33, 1, seed_offset=102)
    make_dense("dense_composite_k9", 64, 9, seed_offset=103)
    make_appearance("appearance_rgb_128", 300, [128, 128, 64], 3, seed=201, normal_dep=False)   # reference default (hyper_params.py:15)
    make_appearance("appearance_alpha_64", 257, [64, 64, 64], 1, seed=202, normal_dep=True)     # "64-wide" (config C4)
