"""Golden vectors for the SH-neural-texture appearance (SURVEY 8a row a6'), produced by IMPORTING THE REFERENCE'S OWN CLASSES
``volsurfs_py.models.sh_neural_textures.SHNeuralTextures`` / ``neural_texture.NeuralTexture`` (and through them
``mvdatasets.utils.images``, ``volsurfs_py.utils.math.round_ste``, ``SHEncoder.eval``) on CPU tensors.

    python tests/golden/make_golden_shtex.py        (in the container that mounts /root/reference)

The only thing that cannot be imported is tiny-cuda-nn (un-vendored, GPU only): ``tinycudann`` is replaced by a stub whose
``Encoding`` / ``Network`` call the restatement in oracle/shtex.py (fp16 output).  So these files pin the reference's GLUE bit-exact
(uv -> corners / lerp weights, align_to_webgl, sigmoid, 8-bit STE quantisation, fp16 re-expansion, lerp, coefficient layout, mixed
precision SH evaluation, sigmoid, and autograd through all of it); the tiny-cuda-nn arithmetic stays "parity unpinned".
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(OUT.parent.parent))

from oracle import shtex as O  # noqa: E402


def install_stubs():
    """tinycudann -> restatement; names the reference imports but never uses on this path -> empty modules"""
    tcnn = types.ModuleType("tinycudann")

    class Encoding(torch.nn.Module):
        def __init__(self, n_input_dims, encoding_config):
            super().__init__()
            assert n_input_dims == 2 and encoding_config["otype"] == "HashGrid"
            self.cfg = encoding_config
            self.n_output_dims = encoding_config["n_levels"] * encoding_config["n_features_per_level"]
            self.net = None  # filled by Network (the Sequential is the unit of the restatement)

        def forward(self, x):
            return x  # the paired Network evaluates encoding + MLP

    class Network(torch.nn.Module):
        seeds = iter(range(1000, 2000))

        def __init__(self, n_input_dims, n_output_dims, network_config):
            super().__init__()
            assert network_config["otype"] == "FullyFusedMLP" and network_config["activation"] == "ReLU"
            self.net = O.TextureNet(n_output_dims, seed=next(Network.seeds), n_neurons=network_config["n_neurons"],
                                    n_hidden_layers=network_config["n_hidden_layers"], table_init=0.5)

        def forward(self, uv):
            return self.net(uv)

    tcnn.Encoding, tcnn.Network = Encoding, Network
    sys.modules["tinycudann"] = tcnn
    sys.modules.setdefault("permutohedral_encoding", types.ModuleType("permutohedral_encoding"))
    import scipy.special

    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = scipy.special.sph_harm_y
    sys.path.insert(0, str(REF))
    # mvdatasets/__init__.py pulls in dataset loaders (pycolmap, ...): load the one file the path uses, utils/images.py, directly
    import importlib.util

    for pkg in ("mvdatasets", "mvdatasets.utils"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    spec = importlib.util.spec_from_file_location("mvdatasets.utils.images", REF / "submodules/mvdatasets/mvdatasets/utils/images.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules["mvdatasets.utils.images"] = mod


def make(name: str, n: int, nr_channels: int, sh_deg: int, seed: int, anchor: bool, lerp: bool, deg_res, sh_range):
    from volsurfs_py.models.sh_neural_textures import SHNeuralTextures  # reference class

    model = SHNeuralTextures(sh_deg=sh_deg, nr_channels=nr_channels, sh_range=list(sh_range), anchor=anchor, lerp=lerp,
                             deg_res=list(deg_res), quantize_output=True, squeeze_output=True, align_to_webgl=True)  # volsurfs.py:145-155
    nets = [nt.network.net for nt in model.neural_textures]
    while True:
        g = torch.Generator().manual_seed(seed)
        uv = torch.rand(n, 2, generator=g)
        uv[:8] = torch.tensor([[0.0, 0.0], [1.0, 1.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [1e-4, 0.9999], [0.25, 0.75], [0.999, 0.001]])
        # no ReLU ties: a hidden pre-activation within fp32 summation noise of 0 has an order-dependent derivative
        tie = min(net.min_abs_preactivation(O.texel_queries(uv.clone(), [deg_res[d], deg_res[d]], anchor, lerp, True)[0])
                  for d, net in enumerate(nets))
        if tie > 2e-7:
            break
        print(f"  seed {seed}: smallest hidden pre-activation {tie:.2e} -> next seed")
        seed += 1000
        assert seed < 40000
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    g_out = torch.randn(n, nr_channels, generator=g)
    out = model(uv_coords=uv.clone(), view_dirs=dirs)          # sh_neural_textures.py:64-97
    coeffs = model(uv_coords=uv.clone(), view_dirs=None)       # the coefficient tensor [n, C, nr_coeffs]
    (out * g_out).sum().backward()
    save = {"uv": uv.numpy(), "dirs": dirs.numpy(), "g_out": g_out.numpy(), "out": out.detach().numpy(), "coeffs": coeffs.detach().numpy(),
            "sh_deg": np.array(sh_deg), "nr_channels": np.array(nr_channels), "anchor": np.array(anchor), "lerp": np.array(lerp),
            "deg_res": np.array(deg_res), "sh_range": np.array(sh_range, np.float32)}
    # the tables are regenerated from their seeds (11 MB each otherwise); a checksum guards against generator drift.  Their gradients
    # are sparse: rows touched by the batch.
    save["seeds"] = np.array([net.seed for net in nets])
    for d, net in enumerate(nets):
        save[f"table{d}_sum"] = np.array(net.table.detach().double().sum().item())
        gt = net.table.grad
        rows = torch.nonzero(gt.abs().sum(1) > 0).flatten()
        save[f"d_table{d}_rows"] = rows.numpy()
        save[f"d_table{d}_vals"] = gt[rows].numpy()
        for i, W in enumerate(net.weights):
            save[f"W{d}_{i}"] = W.detach().numpy()
            save[f"dW{d}_{i}"] = W.grad.numpy()
    np.savez_compressed(OUT / f"{name}.npz", **save)
    print("wrote", name, out.shape, coeffs.shape, float(out.mean()))


if __name__ == "__main__":
    assert REF.exists(), "run in the container that mounts /root/reference"
    install_stubs()
    # small texture resolutions keep the fixtures small; the table init is widened (0.5 instead of tcnn's 1e-4) so that the
    # outputs span the sigmoid / quantisation range
    make("shtex_rgb_lerp", 200, 3, 3, seed=301, anchor=False, lerp=True, deg_res=[64, 32, 16, 8], sh_range=[15.0, 15.0, 15.0, 15.0])
    make("shtex_alpha_lerp", 160, 1, 3, seed=302, anchor=False, lerp=True, deg_res=[2048, 1024, 512, 256], sh_range=[1.0, 5.0, 10.0, 20.0])
    make("shtex_rgb_anchor", 120, 3, 2, seed=303, anchor=True, lerp=False, deg_res=[64, 32, 16, 8], sh_range=[15.0, 15.0, 15.0, 15.0])
