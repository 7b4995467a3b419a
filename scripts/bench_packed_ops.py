"""Times the packed VolumeRendering forward operators through the PyBridge-shaped shim against the
reference's own kernels recompiled for sm_100a (oracle/_ref/libvolsurfs_ref.so), on BASELINE config[2]-shaped packets
(occupancy-grid ray marching, up to 1024 samples per ray; arrays larger than L2).
    python scripts/bench_packed_ops.py [n_rays] [reps]
Product arm: CUDA events around the public call (output allocation + fill included, as a caller sees it).  Reference arm: its kernel
alone on preallocated outputs (the reference's host wrapper additionally allocates, fills and cudaDeviceSynchronize()s — not timed, so
the comparison is conservative).  GB/s = algorithmic bytes (inputs read once + outputs written once) / median time.
Checker-only use of oracle/_ref (measurement tool, not product code).
"""
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from volsurfs_b200.synthetic import nerf_packets  # noqa: E402
from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 11
p = nerf_packets(n, seed_offset=3)
S = int(p["alpha"].shape[0])
rsp = RaySamplesPacked(0, 0, 0, 1)
rsp.ray_start_end_idx = p["se"].cuda()
rsp.samples_z = p["z"].cuda()
rsp.samples_dt = p["dt"].cuda()
rsp.has_dt = True
se = rsp.ray_start_end_idx
g = torch.Generator(device="cuda").manual_seed(1)
x = p["x"].cuda()
v1, v3 = torch.randn(S, 1, device="cuda", generator=g), torch.randn(S, 3, device="cuda", generator=g)
w = torch.rand(S, 1, device="cuda", generator=g) / 64
sdf, beta = torch.randn(S, 1, device="cuda", generator=g) * 0.05, torch.full((S, 1), 200.0, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
ref_path = ROOT / "oracle" / "_ref" / "libvolsurfs_ref.so"
ref = ctypes.CDLL(str(ref_path)) if ref_path.exists() else None
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else None
o_s1, o_s1b, o_n1, o_n3 = torch.zeros(S, 1, device="cuda"), torch.zeros(S, 1, device="cuda"), torch.zeros(n, 1, device="cuda"), torch.zeros(n, 3, device="cuda")


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


def ok(code):
    assert code == 0, code


# name, algorithmic bytes, product call, reference-kernel call
OPS = [
    ("cumprod_one_minus_alpha_to_transmittance", 8 * S + 12 * n, lambda: VR.cumprod_one_minus_alpha_to_transmittance(rsp, x),
     lambda: ok(ref.ref_cumprod_fwd(P(se), P(x), P(o_s1), P(o_n1), n, S))),
    ("integrate_with_weights_1d", 8 * S + 12 * n, lambda: VR.integrate_with_weights_1d(rsp, v1, w),
     lambda: ok(ref.ref_integrate_fwd(P(se), P(v1), P(w), P(o_n1), 1, n, S))),
    ("integrate_with_weights_3d", 16 * S + 20 * n, lambda: VR.integrate_with_weights_3d(rsp, v3, w),
     lambda: ok(ref.ref_integrate_fwd(P(se), P(v3), P(w), P(o_n3), 3, n, S))),
    ("sum_over_rays(d=1)", 8 * S + 12 * n, lambda: VR.sum_over_rays(rsp, v1),
     lambda: ok(ref.ref_sum_fwd(P(se), P(v1), P(o_n1), P(o_s1), 1, n, S))),
    ("cumsum_over_rays", 8 * S + 8 * n, lambda: VR.cumsum_over_rays(rsp, v1, False),
     lambda: ok(ref.ref_cumsum(P(se), P(v1), 0, P(o_s1), n, S))),
    ("sdf2alpha", 16 * S + 8 * n, lambda: VR.sdf2alpha(rsp, sdf, beta),
     lambda: ok(ref.ref_sdf2alpha(P(se), P(rsp.samples_dt), P(sdf), P(beta), P(o_s1), n, S))),
    ("compute_cdf", 8 * S + 8 * n, lambda: VR.compute_cdf(rsp, w),
     lambda: ok(ref.ref_compute_cdf(P(se), P(w), P(o_s1b), n, S))),
    ("median_depth_over_rays", 8 * S + 12 * n, lambda: VR.median_depth_over_rays(rsp, w, 0.5),
     lambda: ok(ref.ref_median_depth(P(se), P(rsp.samples_z), P(w), ctypes.c_float(0.5), P(o_n1), n, S))),
]
print(json.dumps({"n_rays": n, "n_samples": S, "hbm_peak_gbs": peak, "reps": reps}))
for name, nbytes, prod, refk in OPS:
    row = {"op": name, "algorithmic_MB": round(nbytes / 1e6, 1)}
    try:
        ms = timed(prod)
        row["product_ms"], row["product_gbs"] = round(ms, 4), round(nbytes / ms / 1e6, 1)
        if peak:
            row["product_frac_of_peak"] = round(nbytes / ms / 1e6 / peak, 3)
        if ref is not None:
            ms_r = timed(refk)
            row["reference_kernel_ms"], row["speedup"] = round(ms_r, 4), round(ms_r / ms, 2)
        if name == "median_depth_over_rays":
            # like for like with the reference arm (its kernel alone): the product's entry point on a preallocated, pre-zeroed output --
            # on a 40 us operator the public call's allocation + zero fill is a quarter of the time
            from volsurfs_b200 import _lib
            o_k = torch.zeros(n, 1, device="cuda")
            st = torch.cuda.current_stream().cuda_stream
            ms_k = timed(lambda: _lib.check(_lib.lib().vs_median_depth(se.data_ptr(), rsp.samples_z.data_ptr(), w.data_ptr(), 0.5,
                                                                         o_k.data_ptr(), n, S, 0, st), "vs_median_depth"))
            row["product_kernel_ms"] = round(ms_k, 4)
            if ref is not None:
                row["speedup_kernel_only"] = round(ms_r / ms_k, 2)
    except Exception as e:  # one op failing must not hide the others
        row["error"] = repr(e)[:200]
    print(json.dumps(row), flush=True)
