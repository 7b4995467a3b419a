"""CPU: the SH-neural-texture restatement (oracle/shtex.py) against the vectors produced by the reference's own SHNeuralTextures /
NeuralTexture classes (tests/golden/make_golden_shtex.py; tiny-cuda-nn stubbed by the restatement) — bit-exact, outputs and autograd
gradients; plus known-answer checks of the tiny-cuda-nn grid geometry."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import shtex as O

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["shtex_rgb_lerp", "shtex_alpha_lerp", "shtex_rgb_anchor"]


def load_case(name):
    d = np.load(GOLDEN / f"{name}.npz")
    C, deg = int(d["nr_channels"]), int(d["sh_deg"])
    nets = [O.TextureNet(C * O.DEG_NR_COEFFS[g], seed=int(d["seeds"][g]), table_init=0.5) for g in range(deg + 1)]
    for g, net in enumerate(nets):
        assert abs(float(net.table.detach().double().sum()) - float(d[f"table{g}_sum"])) < 1e-9, "generator drift: regenerate the goldens"
        for i, W in enumerate(net.weights):
            assert np.array_equal(W.detach().numpy(), d[f"W{g}_{i}"])
    return d, nets, C, deg


def run_oracle(d, nets, C, deg, with_dirs=True):
    return O.sh_neural_textures_forward(
        nets, torch.from_numpy(d["uv"]).clone(), torch.from_numpy(d["dirs"]) if with_dirs else None, sh_deg=deg, nr_channels=C,
        sh_range=[float(v) for v in d["sh_range"]], deg_res=[int(v) for v in d["deg_res"]], anchor=bool(d["anchor"]), lerp=bool(d["lerp"]))


@pytest.mark.parametrize("name", CASES)
def test_restatement_matches_reference_classes(name):
    d, nets, C, deg = load_case(name)
    coeffs = run_oracle(d, nets, C, deg, with_dirs=False)
    assert np.array_equal(coeffs.detach().numpy(), d["coeffs"])
    out = run_oracle(d, nets, C, deg)
    assert np.array_equal(out.detach().numpy(), d["out"])
    (out * torch.from_numpy(d["g_out"])).sum().backward()
    for g, net in enumerate(nets):
        rows = torch.from_numpy(d[f"d_table{g}_rows"])
        gt = net.table.grad
        assert np.array_equal(gt[rows].numpy(), d[f"d_table{g}_vals"])
        mask = torch.ones(gt.shape[0], dtype=torch.bool)
        mask[rows] = False
        assert float(gt[mask].abs().max()) == 0.0
        for i, W in enumerate(net.weights):
            assert np.array_equal(W.grad.numpy(), d[f"dW{g}_{i}"])


def test_hashgrid_geometry_known_answers():
    """tiny-cuda-nn grid.h for NeuralTexture's config (neural_texture.py:54-61): base 16, scale 1.5, 2^15 entries, 16 levels"""
    levels, total = O.hashgrid_levels()
    assert [l["res"] for l in levels[:7]] == [16, 24, 36, 54, 81, 122, 183]
    assert [l["size"] for l in levels[:7]] == [256, 576, 1296, 2920, 6568, 14888, 32768]
    assert [l["hashed"] for l in levels] == [False] * 6 + [True] * 10
    assert total == 256 + 576 + 1296 + 2920 + 6568 + 14888 + 10 * 32768
    # dense level: corner 0 of x = (0,0) is entry 0 with weight (1-0.5)^2; hashed level: index = x ^ y*2654435761 mod 2^15
    idx, w = O.hashgrid_corner_terms(np.zeros((1, 2), np.float32), levels[0])
    assert idx[0].tolist() == [0, 1, 16, 17] and np.allclose(w[0], 0.25)
    idx, _ = O.hashgrid_corner_terms(np.array([[0.5, 0.25]], np.float32), levels[8])
    s = float(levels[8]["scale"])
    gx, gy = int(np.floor(np.float32(0.5) * np.float32(s) + np.float32(0.5))), int(np.floor(np.float32(0.25) * np.float32(s) + np.float32(0.5)))
    assert idx[0, 0] == ((gx ^ ((gy * 2654435761) & 0xFFFFFFFF)) % 32768)


def test_quantisation_levels_and_range():
    """squeeze + quantise + fp16 expansion (neural_texture.py:155-181): coefficients sit on the 256-level lattice of val_range"""
    d, nets, C, deg = load_case("shtex_rgb_anchor")
    coeffs = run_oracle(d, nets, C, deg, with_dirs=False).detach().numpy()
    r = float(d["sh_range"][0])
    assert np.all(np.abs(coeffs) <= r + 1e-3)
    q = (coeffs[:, :, 0] + r) / (2 * r) * 255.0
    assert np.max(np.abs(q - np.round(q))) < 0.2  # fp16 expansion error only (three fp16 roundings of values up to 15)


def test_c_abi_level_table_matches_restatement(lib):
    """vs_hashgrid_levels runs on the host: the kernels' level geometry == the restatement's, scale bit-equal"""
    import ctypes

    L = 16
    scale = (ctypes.c_float * L)()
    res, size, off = (ctypes.c_int32 * L)(), (ctypes.c_int32 * L)(), (ctypes.c_int32 * L)()
    total = lib.vs_hashgrid_levels(L, 15, 16, ctypes.c_float(1.5), scale, res, size, off)
    levels, want_total = O.hashgrid_levels()
    assert total == want_total
    assert [np.float32(s) for s in scale] == [l["scale"] for l in levels]
    assert list(res) == [l["res"] for l in levels] and list(size) == [l["size"] for l in levels] and list(off) == [l["offset"] for l in levels]
    assert lib.vs_hashgrid_levels(17, 15, 16, ctypes.c_float(1.5), None, None, None, None) < 0
