"""Times the occupancy-grid foreground sampler (the producer of BASELINE config C3's packets) against the reference's
own kernel recompiled for sm_100a (oracle/_ref/libsampler_ref.so) followed by the compaction the reference runs afterwards.
    python scripts/bench_sampler.py [n_rays] [voxels_per_dim] [max_samples_per_ray] [reps]
Both arms: CUDA events around the public call (allocation of the outputs included, as in the reference's RaySampler); the reference arm
also reports its kernel alone.  Checker-only use of oracle/_ref (this script is a measurement tool, not product code).
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from sampler_scene import make_scene  # noqa: E402
from volsurfs_b200.volsurfs import RaySampler, RaySamplesPacked  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 640000
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 128
max_nr = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
min_dist, min_nr = 0.0015, 1
sc = make_scene(n, nv, seed=31)
t = {k: torch.from_numpy(np.ascontiguousarray(sc[k])).cuda() for k in ("o", "d", "t_entry", "t_exit", "occ", "roi")}
P = lambda x: ctypes.c_void_p(x.data_ptr())  # noqa: E731
med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731


def product():
    return RaySampler.compute_samples_fg_in_grid_occupied_regions(t["o"], t["d"], t["t_entry"], t["t_exit"], min_dist, min_nr, max_nr, False, nv,
                                                                  sc["extent"], t["occ"], t["roi"], 1)


ref_path = ROOT / "oracle" / "_ref" / "libsampler_ref.so"
ref = ctypes.CDLL(str(ref_path)) if ref_path.exists() else None
ext = (ctypes.c_float * 3)(*[float(v) for v in sc["extent"]])
kernel_ms = []


def reference():
    unc = RaySamplesPacked(n, n * max_nr, 0, 1)
    unc.is_compacted = False
    unc.ray_o, unc.ray_d, unc.ray_enter, unc.ray_exit = t["o"].clone(), t["d"].clone(), t["t_entry"].clone(), t["t_exit"].clone()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    code = ref.ref_samples_fg_occupied(P(t["o"]), P(t["d"]), P(t["t_entry"]), P(t["t_exit"]), ctypes.c_float(min_dist), min_nr, max_nr,
                                       ctypes.c_uint64(0x853C49E6748FEA9B), ctypes.c_uint64(0xDA3E39CB94B95BDB), 0, nv, ext, P(t["occ"]), P(t["roi"]),
                                       P(unc.ray_max_dt), P(unc.samples_idx), P(unc.samples_3d), P(unc.samples_dirs), P(unc.samples_z),
                                       P(unc.samples_dt), P(unc.ray_start_end_idx), n)
    k1.record()
    assert code == 0
    out = unc.compact_to_valid_samples()
    torch.cuda.synchronize()
    kernel_ms.append(k0.elapsed_time(k1))
    return out


def timeit(fn):
    ts, out = [], None
    for it in range(reps + 1):
        out = None
        torch.cuda.empty_cache()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 1:
            ts.append(e0.elapsed_time(e1))
    return med(ts), out


ms_p, out_p = timeit(product)
S = out_p.get_total_nr_samples()
line = (f"sampler fg-in-grid n_rays={n} grid={nv}^3 max/ray={max_nr}: {S} samples ({S / n:.1f}/ray) | product {ms_p:.3f} ms = "
        f"{n / ms_p / 1e3:.1f} Mrays/s, {S / ms_p / 1e3:.1f} Msamples/s, {S * 32 / ms_p / 1e6:.0f} GB/s of sample rows")
if ref is not None:
    z_p, se_p = out_p.samples_z.clone(), out_p.ray_start_end_idx.clone()
    del out_p
    ms_r, out_r = timeit(reference)
    same = torch.equal(out_r.samples_z, z_p) and torch.equal(out_r.ray_start_end_idx, se_p)
    line += (f" | reference kernel (sm_100a build) + compaction {ms_r:.3f} ms (kernel alone {med(kernel_ms[1:]):.3f} ms) -> x{ms_r / ms_p:.2f}; "
             f"outputs identical: {same}")
print(line, flush=True)
