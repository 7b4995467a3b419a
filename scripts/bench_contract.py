"""Times the background scene contraction (RaySampler.contract_samples' kernel, src/RaySampler.cu:336-381) against the reference's own
kernel recompiled for sm_100a (oracle/_ref/libsampler_ref.so), on a background packet of n_rays x nr samples (inputs > L2).
    python scripts/bench_contract.py [n_rays] [samples_per_ray] [reps]
CUDA events on torch's current stream (the stream the C ABI launches on); the reference kernel runs on the legacy default stream, which
is torch's current stream here.  Algorithmic bytes: 16 B in + 16 B out per sample + 20 B per ray.  Checker-only use of oracle/_ref.
"""
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from volsurfs_b200 import _lib  # noqa: E402
from volsurfs_b200.volsurfs import RaySampler  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
g = torch.Generator(device="cuda").manual_seed(3)
o = (torch.rand(n, 3, device="cuda", generator=g) - 0.5) * 0.4
d = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=1).contiguous()
rsp = RaySampler.compute_samples_bg(o, d, torch.full((n, 1), 0.4, device="cuda"), 40.0, nr, True)
out3, outz = torch.empty_like(rsp.samples_3d), torch.empty_like(rsp.samples_z)
P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
L = _lib.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ref_path = ROOT / "oracle" / "_ref" / "libsampler_ref.so"
ref = ctypes.CDLL(str(ref_path)) if ref_path.exists() else None


def product():
    assert L.vs_sampler_contract(P(rsp.ray_o), P(rsp.ray_start_end_idx), P(rsp.samples_3d), P(rsp.samples_z), P(out3), P(outz), 0, n, st) == 0


def reference():
    assert ref.ref_contract_samples(P(rsp.ray_o), P(rsp.ray_start_end_idx), P(rsp.samples_3d), P(rsp.samples_z), P(out3), P(outz), n, n * nr, 0) == 0


def timed(fn):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2]


peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else None
nbytes = n * nr * 32 + n * 20
row = {"n_rays": n, "samples_per_ray": nr, "algorithmic_MB": round(nbytes / 1e6, 1), "hbm_peak_gbs": peak}
ms = timed(product)
row["product_ms"], row["product_gbs"] = round(ms, 4), round(nbytes / ms / 1e6, 1)
if peak:
    row["product_frac_of_peak"] = round(nbytes / ms / 1e6 / peak, 3)
if ref is not None:
    ms = timed(reference)
    row["reference_kernel_ms"], row["reference_kernel_gbs"] = round(ms, 4), round(nbytes / ms / 1e6, 1)
print(json.dumps(row))
