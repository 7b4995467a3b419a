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


if __name__ == "__main__":
    assert REF.exists(), "run in the container that mounts /root/reference"
    make_dense("dense_composite_k5", 96, 5, seed_offset=101)
    make_dense("dense_composite_k1", 33, 1, seed_offset=102)
    make_dense("dense_composite_k9", 64, 9, seed_offset=103)
