"""Times the SH-neural-texture appearance (SURVEY 8a row a6') stage by stage for one SHNeuralTextures model (CUDA events, L2 flushed).
    python scripts/bench_shtex.py [n_hits] [nr_channels] [reps]
Per hit and SH degree g (lerp mode): 4 texel queries x (16 levels x 4 gathers of 8 B from an L2-resident table) -> 4 x 128 B of feature
rows -> 4 x 2*(32*64 + 64*64 + 64*pad16(C*(2g+1))) tensor-core flop -> 4 x C*(2g+1) raw outputs; then one combine kernel per model.
"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200 import _lib  # noqa: E402
from volsurfs_b200.textures import SHNeuralTextures, _combine_args  # noqa: E402
from volsurfs_b200.volsurfs import _stream  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 178548  # hits of ONE layer at BASELINE config[1] (892741 / 5)
C = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
torch.manual_seed(0)
m = SHNeuralTextures(sh_deg=3, nr_channels=C, sh_range=[15.0] * 4, lerp=True, deg_res=[2048, 1024, 512, 256], quantize_output=True,
                     squeeze_output=True, align_to_webgl=True).cuda()
with torch.no_grad():
    for nt in m.neural_textures:
        nt.model.table.copy_((torch.rand_like(nt.model.table) * 2 - 1) * 0.5)
nets = [nt.model for nt in m.neural_textures]
uv = torch.rand(n, 2, device="cuda")
dirs = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda"), dim=1)
g_out = torch.randn(n, C, device="cuda")
L = _lib.lib()
res = [[r, r] for r in (2048, 1024, 512, 256)]
args = _combine_args(3, C, 1, True, res, [(-15.0, 15.0)] * 4, True, True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["encode", "mlp_fwd", "combine_fwd", "combine_bwd", "mlp_bwd+hashgrid_bwd"]
acc = {k: [] for k in names}
stashes = [net.new_stash(4 * n, "cuda") for net in nets]
for it in range(reps + 2):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    feats = [net.encode(uv, 1, True, res[g]) for g, net in enumerate(nets)]
    ev[1].record()
    raws = [net.mlp_raw(feats[g], stashes[g]) for g, net in enumerate(nets)]
    ev[2].record()
    out = torch.empty(n, C, device="cuda")
    rp = (ctypes.c_void_p * 4)(*[r.data_ptr() for r in raws])
    assert L.vs_shtex_combine_forward(*args, uv.data_ptr(), dirs.data_ptr(), rp, None, out.data_ptr(), n, None, _stream()) == 0
    ev[3].record()
    d_raws = [torch.empty_like(r) for r in raws]
    dp = (ctypes.c_void_p * 4)(*[r.data_ptr() for r in d_raws])
    assert L.vs_shtex_combine_backward(*args, uv.data_ptr(), dirs.data_ptr(), rp, out.data_ptr(), g_out.data_ptr(), None, dp, n, None, _stream()) == 0
    ev[4].record()
    for g, net in enumerate(nets):
        net.backward_into(uv, 1, True, res[g], stashes[g], d_raws[g])
    ev[5].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i, k in enumerate(names):
            acc[k].append(ev[i].elapsed_time(ev[i + 1]))
med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
rows = 4 * n
flops = sum(2 * (32 * 64 + 64 * 64 + 64 * ((C * (2 * g + 1) + 15) // 16 * 16)) for g in range(4)) * rows
t = {k: med(v) for k, v in acc.items()}
total = sum(t.values())
print(f"shtex n_hits={n} C={C} deg=3 lerp ({rows} query rows x 4 networks): " + "  ".join(f"{k} {v:.3f} ms" for k, v in t.items()) +
      f"  | total {total:.3f} ms = {n / total / 1e3:.2f} Mhits/s; encode gathers {rows * 4 * 16 * 4 * 8 / t['encode'] / 1e6:.0f} GB/s + rows out "
      f"{rows * 4 * 128 / t['encode'] / 1e6:.0f} GB/s; mlp_fwd {flops / t['mlp_fwd'] / 1e9:.1f} TFLOP/s", flush=True)
