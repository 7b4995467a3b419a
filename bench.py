#!/usr/bin/env python
"""Benchmark of the per-ray rendering hot path (BASELINE.json metric: Mrays/s fwd+bwd, 5-layer shells; % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic rays on every rank (rays are sharded, weak scaling):
  K-layer shell intersection (one launch) -> hit packing -> face normals -> permutohedral hash encoding of the hit points (one encoder
  per head, as volsurfs_py/models/rgb.py:40-60) -> rgb head + alpha head (tcgen05 MLPs) -> fused compositing forward -> L1 loss gradient
  -> fused compositing backward (d_alpha, d_rgb per hit) -> backward of both heads (tcgen05: Linear weight/bias gradients + gradients of
  the positional features) -> backward of both encoders (lattice gradients); (N > 1) ONE NCCL all-reduce (mean) of the head AND lattice
  gradients (100.8 MB, one flat buffer), captured inside the step's CUDA graph on a side stream under the NEXT step's trace + pack
  (software pipelining by one step: every replay performs one whole step and one whole exchange).
Workload at every N: BASELINE config[1] per GPU — 800x800 camera rays against 5 nested ~100k-triangle shells, legacy
[128,128,64] GELU heads on the 51 features of a 24-level x 2, 2^18-entry permutohedral encoder + SH deg 3.

Prints ONE JSON line (rank 0).  `value` is the device-resident whole-job throughput, `e2e` the same step driven from pinned
host ray buffers with the image copied back, `roofline` describes the dominant kernel of the step, `stages` every stage,
`compositing_roofline` the headline compositing kernels at 2^24 rays (the size SURVEY.md section 8d prescribes),
`cpu_baseline` the reference algorithm (oracle port: C BVH tracer + torch CPU heads + the reference's dense torch
compositing differentiated by autograd) on a bounded sample on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mrays/s fwd+bwd, 5-layer shells"
UNIT = "Mrays/s"
K_LAYERS = 5
IMG = 800
HIDDEN = (128, 128, 64)
POS_DIM = 51
CPU_SAMPLE_RAYS = 16384  # the reference's own render chunk (config/volsurfs/base_5.cfg:8)
# dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full captures of this workload at HEAD (profiles/r02_*.md)
NCU_TRAFFIC = {"mlp_fwd_kernel": 204.320000e6 + 693.690368e6,            # profiles/r02_mlp_fwd.md
               "mlp_bwd_stashed_kernel": 764.562432e6 + 169.221120e6,    # profiles/r02_mlp_bwd.md
               "shells_trace_kernel": 44.94848e6 + 19.424512e6,         # profiles/r02_shells_trace.md
               "permuto_fwd_kernel": 31.724800e6 + 133.739264e6,         # profiles/r02_permuto_fwd.md
               "permuto_bwd_kernel": 214.678784e6 + 8.321536e6,          # profiles/r02_permuto_bwd.md
               # composite_fwd_tile_kernel<1> + composite_bwd_tile_kernel<1> at 2^24 rays x 5 (algorithmic: 5.77 GB)
               "composite_tile_fwd+bwd": 1.812154e9 + 389.939712e6 + 2.214624e9 + 1.307345e9}   # profiles/r02_composite_{fwd,bwd}_tile.md


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything libraries print on fd 1 (NCCL's version banner, OpenMP notices) goes to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml; nvidia-smi fallback)"""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _sample(self):
        if self.nv is None:
            return
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for name, bit in {**self.BAD, **self.NOTE}.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:  # noqa: BLE001
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.02)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._sample()
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": int(statistics.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's algorithm on the host cores (oracle port)
# ----------------------------------------------------------------------------------------------------------------------
class CpuReferencePath:
    """The reference's CPU-runnable path, restated by oracle/: per-mesh BVH trace (C, OpenMP, raytracelib algorithm), per-mesh
    shading of the hit points with the legacy torch heads, the dense [N,K] torch compositing of volsurfs.py:601-640,708 and its
    autograd backward down to the per-surface colour/alpha (the same depth our step differentiates to)."""

    def __init__(self, n_rays: int):
        import numpy as np
        import torch

        from oracle import appearance as oa
        from oracle.raytrace import OracleRayTracer
        from volsurfs_b200.synthetic import camera_rays, shell_meshes

        self.torch, self.np, self.oa = torch, np, oa
        torch.set_num_threads(os.cpu_count() or 1)
        self.meshes = shell_meshes(K=K_LAYERS)
        self.tracer = OracleRayTracer(self.meshes)
        o, d = camera_rays(IMG, IMG)
        # a contiguous block of rows through the middle of the image (same hit density as the whole frame's centre)
        first = (IMG // 2) * IMG - n_rays // 2
        self.o, self.d = o[first:first + n_rays].contiguous(), d[first:first + n_rays].contiguous()
        self.n_rays = n_rays
        self.rgb_w = oa.init_linear_stack(POS_DIM + 16, HIDDEN, 3, seed=11)
        self.alpha_w = oa.init_linear_stack(POS_DIM + 16, HIDDEN, 1, seed=12)
        g = torch.Generator().manual_seed(13)
        from oracle import permuto as op

        self.op = op
        # one permutohedral encoder per head (rgb.py:40-60; encodings/permutohash.py:10-41): 24 levels x 2, 2^18 entries, concat points
        self.encs = [op.PermutoEncoding(3, 2 ** 18, 24, 2, np.geomspace(1.0, 1e-4, 24), True, True, 1.0, seed=21 + i) for i in range(2)]
        self.lattice_grads = [np.zeros_like(e.lattice_values) for e in self.encs]
        for stack in (self.rgb_w, self.alpha_w):
            for lst in stack:
                for t in lst:
                    t.requires_grad_(True)
        self.gt = torch.rand(n_rays, 3, generator=g)
        self.cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(self.cores))

    def step(self):
        torch, np, oa = self.torch, self.np, self.oa
        from oracle.compositing import dense_composite_torch

        N, K = self.n_rays, K_LAYERS
        o, d = self.o.numpy(), self.d.numpy()
        op = self.op
        surfs_rgb = torch.zeros(N, K, 3)
        surfs_alpha = torch.zeros(N, K, 1)
        leaves = []
        for i in range(K):                                        # volsurfs.py:476-485
            res = self.tracer.trace(o, d, i, mode="bvh")
            if not res["any_hit"]:
                continue
            hits = torch.from_numpy(res["is_hit"])
            normals = torch.from_numpy(res["normals"])[hits]
            dirs = self.d[hits]
            unit, _ = op.volsurfs_points_to_unit_cube(res["positions"][res["is_hit"]], 2.0)   # permutohash.py:77-86
            feats = []
            for e in self.encs:                                   # encoder forward (C twin of the numpy oracle, all host threads)
                rows = op.forward_rows_c(unit, e.lattice_values, e.scale, e.random_shift_per_level, e.anneal_window, True, 1.0)
                feats.append(torch.from_numpy(rows[:, :POS_DIM].copy()).requires_grad_(True))   # permutohash.py:91-94: last column dropped
            leaves.append((unit, feats))
            rgb = oa.head_forward(feats[0], dirs, normals, *self.rgb_w)
            alpha = oa.alpha_decay(oa.head_forward(feats[1], dirs, normals, *self.alpha_w), dirs, normals)
            surfs_rgb = surfs_rgb.index_put((hits, torch.tensor(i)), rgb)
            surfs_alpha = surfs_alpha.index_put((hits, torch.tensor(i)), alpha)
        out = dense_composite_torch(surfs_alpha, surfs_rgb, rgb_bg=torch.ones(N, 3))   # volsurfs.py:601-640,708 (fp32 variant)
        loss = (out["rgb"] - self.gt).abs().mean()                                       # utils/losses.py:14-19
        for t in [*self.rgb_w[0], *self.rgb_w[1], *self.alpha_w[0], *self.alpha_w[1]]:
            t.grad = None
        loss.backward()                                           # autograd: compositing + both heads (weights, biases, features)
        for t in self.lattice_grads:
            t.fill(0.0)
        for unit, feats in leaves:                                # encoder backward: feature gradients -> lattice gradients
            for k, e in enumerate(self.encs):
                g = np.zeros((unit.shape[0], 2 * 26), np.float32)
                g[:, :POS_DIM] = feats[k].grad.numpy()
                op.backward_lattice_c(unit, e.lattice_values.shape, e.scale, e.random_shift_per_level, e.anneal_window, g,
                                      out=self.lattice_grads[k])
        return float(loss.detach())


def cpu_compositing_c1(warmup: int = 20, iters: int = 200):
    """SURVEY 8(d)'s CPU baseline protocol: the reference's torch compositing path (volsurfs_py/methods/volsurfs.py:601-640,708, restated
    by oracle/compositing.py:dense_composite_torch and pinned bit-exact to those very lines by tests/golden/dense_composite_*.npz) on CPU
    tensors for BASELINE config[0] (4096 rays x 5 layers), forward + autograd backward, fp32 and the reference-faithful fp16 variant,
    all physical cores, 20 warm-ups + 200 timed iterations, median."""
    import torch

    from oracle.compositing import dense_composite_torch
    from volsurfs_b200.synthetic import dense_layers

    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    n_log = os.cpu_count() or 1
    try:
        import psutil

        n_phys = psutil.cpu_count(logical=False) or n_log
    except Exception:  # noqa: BLE001
        n_phys = n_log
    torch.set_num_threads(n_phys)
    dl = dense_layers(4096, 5, seed_offset=1)
    out = {"workload": "BASELINE config[0]: dense K-layer compositing fwd+bwd, 4096 rays x 5 layers, CPU tensors", "cpu_model": model,
           "physical_cores": n_phys, "logical_cpus": n_log, "torch_threads": torch.get_num_threads(), "warmup": warmup, "iters": iters}
    for half in (False, True):
        a = dl["alpha"].clone().requires_grad_(True)
        c = dl["rgb"].clone().requires_grad_(True)
        bg = torch.ones(4096, 3)
        ts = []
        for it in range(warmup + iters):
            a.grad = c.grad = None
            t0 = time.perf_counter()
            o = dense_composite_torch(a, c, rgb_bg=bg, half=half)
            ((o["rgb"] * dl["g_rgb"]).sum() + (o["bg_transmittance"] * dl["g_bgT"]).sum()).backward()
            if it >= warmup:
                ts.append(time.perf_counter() - t0)
        ms = statistics.median(ts) * 1e3
        key = "fp16_reference_faithful" if half else "fp32"
        out[key] = {"ms_median": round(ms, 4), "mrays_s": round(4096 / ms / 1e3, 3)}
    return out


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    path = CpuReferencePath(CPU_SAMPLE_RAYS)
    for _ in range(max(args.warmup, 1)):
        path.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        path.step()
    dt = time.perf_counter() - t0
    value = path.n_rays * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world) | {"sample": f"{path.n_rays} rays per step (reference render chunk) of the 800x800 frame"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": path.cores, "kind": "port",
                         "sample": f"{args.steps} steps x {path.n_rays} rays (rows through the image centre), all {path.cores} host threads"},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    return {
        "workload": f"BASELINE config[1]: 5-mesh volsurf render of kitten-shaped synthetic shells (~100k tris/layer), {IMG}x{IMG} rays per GPU",
        "rays_per_gpu": IMG * IMG, "layers": K_LAYERS, "triangles_per_layer": 99904, "heads": f"rgb+alpha legacy MLP {list(HIDDEN)} GELU, "
        f"{POS_DIM} positional features of a permutohedral hash encoder per head (24 levels x 2, 2^18 entries) + SH deg 3",
        "step": "trace(K layers, 1 launch) + pack + normals + 2 encoders fwd + 2 MLP heads fwd + composite fwd + L1 grad + composite bwd + "
                "2 MLP heads bwd (dW, db, d_features) + 2 encoders bwd (lattice gradients)"
                + (" + NCCL all-reduce (mean) of the head and lattice gradients inside the step's graph" if world > 1 else ""),
        "parallelism": f"rays sharded over {world} GPU(s) (every rank renders the same 800x800 view: equal hit counts); no data-path "
                       "collective in rendering, one gradient all-reduce per step in training (100.8 MB of fp32 lattice + head gradients)",
        "l2": "inputs_larger_than_l2 (2 x 182 MB features, 2 x 743 MB activation stash, 200 MB packed arrays, 2 x 50 MB lattices per step)",
    }


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from volsurfs_b200 import _lib
    from volsurfs_b200.pipeline import make_synthetic_renderer
    from volsurfs_b200.synthetic import all_hit_packed, camera_rays, composite_bytes, nerf_packets
    from volsurfs_b200.volsurfs import RaySamplesPacked, VolumeRendering as VR

    assert torch.cuda.is_available(), "bench.py needs a CUDA device for the default arm (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _lib.lib()
    pk = peaks()

    renderer, _ = make_synthetic_renderer(K=K_LAYERS, hidden=HIDDEN, pos_dim=POS_DIM)
    # every rank renders the same view (equal hit counts: the max-over-ranks step time then shows the exchange, not view imbalance)
    o_h, d_h = camera_rays(IMG, IMG, azimuth_deg=30.0)
    N = o_h.shape[0]
    o_pin, d_pin = o_h.pin_memory(), d_h.pin_memory()
    rays_o, rays_d = o_h.to(dev), d_h.to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    gt = torch.rand(N, 3, generator=g).to(dev)
    img_pin = torch.empty((N, 3), dtype=torch.float32).pin_memory()
    loss_pin = torch.empty((), dtype=torch.float32).pin_memory()

    from volsurfs_b200.encoding import PermutoHashEncoder

    torch.manual_seed(4321)  # identical initial parameters on every rank
    encs = {"rgb": PermutoHashEncoder(bb_sides=2.0, device=dev), "alpha": PermutoHashEncoder(bb_sides=2.0, device=dev)}
    for e in encs.values():  # the reference starts the lattice at N(0, 1e-5) (modules.py:39-41): features ~0; give the heads a signal
        with torch.no_grad():
            e.encoder.lattice_values.normal_(0.0, 0.1)
    assert encs["rgb"].output_dim == POS_DIM
    heads = {"rgb": renderer.rgb_head, "alpha": renderer.alpha_head}
    S_cap = N * K_LAYERS

    stage_names = ["trace", "pack+normals", "grad_allreduce", "encode_rgb", "encode_alpha", "mlp_rgb", "mlp_alpha", "composite_fwd", "loss_grad",
                   "composite_bwd", "mlp_bwd_rgb", "lattice_bwd_rgb", "mlp_bwd_alpha", "lattice_bwd_alpha"]
    n_marks = len(stage_names) + 1
    feats = {k: torch.zeros((S_cap, POS_DIM), device=dev) for k in heads}
    dfeat = {k: torch.zeros((S_cap, POS_DIM), device=dev) for k in heads}
    # every gradient of the step lives in ONE flat fp32 buffer (lattices first, then the heads), so that the exchange is a single NCCL call
    n_lat = {k: e.encoder.lattice_values.numel() for k, e in encs.items()}
    n_head = {k: h.num_params() for k, h in heads.items()}
    grad_flat = torch.zeros(sum(n_lat.values()) + sum(n_head.values()), device=dev)
    grad_lat, grad_head, off = {}, {}, 0
    for k, e in encs.items():
        grad_lat[k] = grad_flat[off:off + n_lat[k]].view_as(e.encoder.lattice_values)
        off += n_lat[k]
    for k in heads:
        grad_head[k] = grad_flat[off:off + n_head[k]]
        off += n_head[k]
    stash = {k: h.new_stash(S_cap, dev) for k, h in heads.items()}   # activations kept by the training-mode forward for the backward
    allreduce_bytes = grad_flat.numel() * 4
    ar_stream = torch.cuda.Stream(device=dev)

    def all_reduce_grads():
        """ONE mean all-reduce of every gradient of a step (2 x 50.3 MB lattice + 2 x 0.14 MB head) on the current stream.

        Where it runs: software-pipelined by one step.  Intersection and packing do not depend on the parameters, so the exchange of step
        i's gradients runs on a side stream under the trace + pack of step i+1 and is joined before the encoders (where an optimizer step
        would sit) — every replay of the step's graph performs one whole step and one whole exchange.  Overlapping the exchange with the
        BACKWARD of its own step was measured and lost (profiles/r02_scaling_notes.md): NCCL's CTAs under the persistent one-CTA-per-SM head
        kernel delay whole tile columns, under the atomics-bound lattice backward both slow down, and two 50 MB calls cost 0.40 ms where one
        100.7 MB call costs 0.32 ms."""
        dist.all_reduce(grad_flat, op=dist.ReduceOp.AVG)

    @torch.no_grad()  # forward and backward kernels are driven explicitly; no autograd graph in the timed region
    def step(record=None, reduce=True, o=None, d=None):
        from volsurfs_b200.raytracer import pack_layer_hits

        o = rays_o if o is None else o
        d = rays_d if d is None else d

        def mark(i):
            if record is not None:
                record[i].record()

        main = torch.cuda.current_stream()
        mark(0)
        if reduce and world > 1:   # exchange of the PREVIOUS step's gradients, under this step's trace + pack
            ar_stream.wait_stream(main)
            with torch.cuda.stream(ar_stream):
                all_reduce_grads()
        rec = renderer.tracer.trace_layers(o, d)
        mark(1)
        rsp = pack_layer_hits(rec["rays_o"], rec["rays_d"], rec["depth"], rec["tri"], rec["u"], rec["v"], t_far=renderer.tracer.t_far,
                              exact_size=False)
        S = rsp.get_max_nr_samples()
        rsp.samples_normals = torch.empty((S, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.vs_shells_sample_normals(renderer.tracer._handle, rsp.samples_layer.data_ptr(), rsp.samples_triangle.data_ptr(), S,
                                                rsp.total_dev.data_ptr(), rsp.samples_normals.data_ptr(), main.cuda_stream), "normals")
        mark(2)
        if reduce and world > 1:
            main.wait_stream(ar_stream)   # the reduced gradients are complete here: this is where the optimizer step belongs
        mark(3)
        for i, k in enumerate(("rgb", "alpha")):   # permutohash.py:68-96: bounding-box normalisation + lattice slice, one launch
            e = encs[k].encoder
            e._launch_forward(e.lattice_values, rsp.samples_3d, encs[k].window(None), POS_DIM, encs[k].bb_sides, rsp.total_dev, out=feats[k])
            mark(4 + i)
        rgb, _ = heads["rgb"].forward_train(feats["rgb"], rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev, stash=stash["rgb"])
        mark(6)
        alpha, _ = heads["alpha"].forward_train(feats["alpha"], rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev,
                                                stash=stash["alpha"])
        mark(7)
        rgb_fg, depth, acc, bgT = VR.composite(rsp, alpha, rgb, rsp.samples_z)
        mark(8)
        pred, loss, g_pred, g_bgT = renderer.blend_l1(rgb_fg, bgT, gt)   # background blend + L1 loss + its gradient: one launch
        out = {"rgb": pred, "rgb_fg": rgb_fg, "depth": depth, "acc": acc, "bg_transmittance": bgT}
        mark(9)
        d_alpha, d_rgb = renderer.composite_backward(rsp, alpha, rgb, g_pred, g_bgT)
        mark(10)
        fwd_out, d_out = {"rgb": rgb, "alpha": alpha}, {"rgb": d_rgb, "alpha": d_alpha}
        for i, k in enumerate(("rgb", "alpha")):
            heads[k].backward_into(feats[k], rsp.samples_dirs, rsp.samples_normals, d_out[k], grad_head[k], dfeat[k], False, rsp.total_dev,
                                   stash=stash[k], fwd_out=fwd_out[k])
            mark(11 + 2 * i)
            e = encs[k].encoder
            grad_lat[k].zero_()
            e._launch_backward(e.lattice_values, rsp.samples_3d, encs[k].window(None), dfeat[k], encs[k].bb_sides, rsp.total_dev,
                               want_lattice=True, d_lattice=grad_lat[k], order_key=rsp.samples_layer)
            mark(12 + 2 * i)
        out["loss"] = loss
        return out, loss, rsp

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        out, loss, rsp = step()
    torch.cuda.synchronize()
    n_hits = int(rsp.total_dev.item())
    assert not renderer.tracer.overflowed()

    # ---- the step as ONE CUDA graph: ~60 launches of 3-700 us each leave the device waiting for the Python host otherwise.  Stage
    # boundaries are external event-record nodes inside the graph.  With N > 1 the NCCL all-reduces are captured INSIDE the graph (the
    # colour branch's on a forked side stream); if this NCCL / torch build refuses to capture collectives the exchange follows the graph.
    use_graph = not args.no_graph
    graph = graph_e2e = None
    reduce_in_graph = world > 1
    ext = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(n_marks)]
    kernels_per_step = None

    def capture(with_reduce):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            res = step(ext, reduce=with_reduce)
        return gph, res

    if use_graph:
        c0 = lib.vs_launch_count()
        try:
            graph, (g_out, g_loss, g_rsp) = capture(reduce_in_graph)
        except Exception as exc:  # noqa: BLE001
            torch.cuda.synchronize()
            if reduce_in_graph:
                print(f"[bench] capturing the NCCL all-reduce in the graph failed ({exc!r}); exchange after the graph", file=sys.stderr, flush=True)
                reduce_in_graph = False
                try:
                    c0 = lib.vs_launch_count()
                    graph, (g_out, g_loss, g_rsp) = capture(False)
                except Exception as exc2:  # noqa: BLE001
                    print(f"[bench] CUDA graph capture failed ({exc2!r}); running eagerly", file=sys.stderr, flush=True)
                    torch.cuda.synchronize()
                    use_graph, graph = False, None
            else:
                print(f"[bench] CUDA graph capture failed ({exc!r}); running eagerly", file=sys.stderr, flush=True)
                use_graph, graph = False, None
        kernels_per_step = lib.vs_launch_count() - c0 if use_graph else None

    def all_reduce_after_graph():
        if world > 1 and not reduce_in_graph:
            all_reduce_grads()

    def run_step(record=None):
        if use_graph:
            graph.replay()
            all_reduce_after_graph()
        else:
            step(record)

    for _ in range(3):
        run_step()
    torch.cuda.synchronize()
    if use_graph:
        n_hits_graph = int(g_rsp.total_dev.item())
        assert n_hits_graph == n_hits, (n_hits_graph, n_hits)
        assert torch.allclose(g_out["rgb"], out["rgb"]), "graph replay differs from the eager step"

    # ---- timed region: device-resident inputs
    records = [[torch.cuda.Event(enable_timing=True) for _ in range(n_marks)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = lib.vs_launch_count()
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s in range(args.steps):
        run_step(records[s])
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = (kernels_per_step * args.steps) if use_graph else (lib.vs_launch_count() - launches0)
    ms_total = t_start.elapsed_time(t_end)
    if use_graph:
        # stage times: the graph's own event nodes, read after single replays (the nodes are part of every timed replay too)
        per = []
        for _ in range(10):
            graph.replay()
            torch.cuda.synchronize()
            per.append([ext[i].elapsed_time(ext[i + 1]) for i in range(n_marks - 1)])
        stage_ms = [statistics.mean(p[i] for p in per) for i in range(n_marks - 1)]
        if world > 1 and not reduce_in_graph:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            all_reduce_after_graph()
            a1.record()
            torch.cuda.synchronize()
            stage_ms[stage_names.index("grad_allreduce")] = a0.elapsed_time(a1)
    else:
        stage_ms = [statistics.mean(records[s][i].elapsed_time(records[s][i + 1]) for s in range(args.steps)) for i in range(n_marks - 1)]

    ar_alone_ms = 0.0
    if world > 1:
        for _ in range(3):
            all_reduce_grads()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            all_reduce_grads()
        a1.record()
        torch.cuda.synchronize()
        ar_alone_ms = a0.elapsed_time(a1) / 10

    # ---- e2e: pinned host rays in, the step's result out, every step — through the public API (pipeline.PipelinedTrainingStep): two
    # alternating graphs; the copy engines move step i+1's rays in and step i-1's result out while step i computes.  The result of a
    # training step is its loss; --e2e-image also brings the composited image back every step (12 B per ray more: on eight GPUs the box's
    # host memory then serves 184 MB per 3 ms step and becomes the limit, measured 1287 against 1655 Mrays/s device-resident)
    img_back = img_pin if args.e2e_image else None
    e2e_mode = "eager, copies in line with the kernels"
    loop = None
    if use_graph:
        try:
            from volsurfs_b200.pipeline import PipelinedTrainingStep

            loop = PipelinedTrainingStep(renderer, o_pin, d_pin, None, gt, img_back, loss_pin,
                                         step_fn=lambda o, d: step(None, reduce=reduce_in_graph, o=o, d=d)[0])
            e2e_mode = ("PipelinedTrainingStep: 2 alternating CUDA graphs, H2D of the next step's rays and D2H of the previous step's "
                        + ("image + loss" if args.e2e_image else "loss") + " on copy streams forked inside the graph")
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] pipelined e2e capture failed ({exc!r}); eager e2e", file=sys.stderr, flush=True)
            torch.cuda.synchronize()
            loop = None

    def e2e_run(n_steps):
        if loop is not None:
            loop.prime()
            for i in range(n_steps):
                loop.step(i)
                all_reduce_after_graph()
            loop.drain(n_steps)
            return
        for _ in range(n_steps):
            rays_o.copy_(o_pin, non_blocking=True)
            rays_d.copy_(d_pin, non_blocking=True)
            out_e, loss_e, _ = step()
            if args.e2e_image:
                img_pin.copy_(out_e["rgb"], non_blocking=True)
            loss_pin.copy_(loss_e, non_blocking=True)

    e2e_run(3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    torch.cuda.synchronize()
    assert torch.allclose(loss_pin, loss.cpu()), "e2e: the loss read back differs from the eager step"
    if args.e2e_image:
        assert torch.allclose(img_pin, out["rgb"].cpu()), "e2e: the image read back differs from the eager step"

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(t[0]), float(t[1])
        hits_t = torch.tensor([n_hits], device=dev, dtype=torch.int64)
        dist.all_reduce(hits_t)
        n_hits_all = int(hits_t[0])
    else:
        n_hits_all = n_hits
    if rank != 0:
        return

    total_rays = N * world
    value = total_rays * args.steps / (ms_total * 1e-3) / 1e6
    e2e_value = total_rays * args.steps / (ms_e2e * 1e-3) / 1e6

    # ---- per-stage rooflines (rank 0's numbers)
    flops_head = lambda out_dim: 2.0 * n_hits * (67 * 128 + 128 * 128 + 128 * 64 + 64 * out_dim)  # noqa: E731
    # 24 levels x 4 vertices x 8 B gathered (scattered as atomics in the backward) = 768 B per sample, + point in, feature
    # row (or its gradient) out / in.  The lattice traffic is L2-resident (ncu: ~0.25 GB of DRAM per launch), so this is an upper view
    enc_bytes = n_hits * (768 + 12 + POS_DIM * 4)
    stage_alg = {
        "trace": ("hbm", N * K_LAYERS * (24 + 16)),                                   # rays in, (t, face, u, v) out per (ray, layer)
        "pack+normals": ("hbm", N * (16 * K_LAYERS + 8 + 24) + n_hits * (4 + 4 + 12 + 12 + 4 + 4 + 8 + 12)),
        "encode_rgb": ("hbm", enc_bytes), "encode_alpha": ("hbm", enc_bytes),
        "mlp_rgb": ("tensor", flops_head(3)),
        "mlp_alpha": ("tensor", flops_head(1)),
        "composite_fwd": ("hbm", 32 * N + 20 * n_hits),
        "loss_grad": ("hbm", N * 12 * 6),
        "composite_bwd": ("hbm", 32 * N + 36 * n_hits),
        "mlp_bwd_rgb": ("tensor", 2.0 * flops_head(3)),                               # dA + dW GEMMs
        "mlp_bwd_alpha": ("tensor", 2.0 * flops_head(1)),
        "lattice_bwd_rgb": ("hbm", enc_bytes), "lattice_bwd_alpha": ("hbm", enc_bytes),
        "grad_allreduce": ("hbm", 0.0),
    }
    stages = {}
    for name, ms in zip(stage_names, stage_ms):
        bound, alg = stage_alg[name]
        if bound == "hbm":
            ach, peak, unit = alg / (ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
        else:
            ach, peak, unit = alg / (ms * 1e-3) / 1e12, pk["tflops_sustained"], "TFLOP/s"
        stages[name] = {"ms": round(ms, 4), "share": round(ms / sum(stage_ms), 3), "bound": bound, "achieved": round(ach, 2), "peak": peak,
                        "unit": unit, "frac": round(ach / peak, 4)}
    if world > 1:
        stages["grad_allreduce"]["note"] = (
            f"one NCCL all-reduce (mean) of {allreduce_bytes / 1e6:.1f} MB of fp32 gradients (2 lattices + 2 heads, one flat buffer) per step"
            + (", captured inside the step's CUDA graph on a side stream under trace + pack (the previous step's gradients: software "
               "pipelining by one step); this stage is the time the main stream waits at the join before the encoders; the exchange alone: "
               f"{ar_alone_ms:.3f} ms = {allreduce_bytes / 1e9 / (ar_alone_ms * 1e-3):.0f} GB/s algorithm, "
               f"{allreduce_bytes / 1e9 / (ar_alone_ms * 1e-3) * 2 * (world - 1) / world:.0f} GB/s bus bandwidth"
               if reduce_in_graph else " (NOT captured: exchange after the graph)"))
    # the dominant KERNEL of the step: stages that launch the same kernel are summed (both heads run mlp_fwd_kernel / mlp_bwd_stashed_kernel)
    kernel_of = {"trace": "shells_trace_kernel", "mlp_rgb": "mlp_fwd_kernel", "mlp_alpha": "mlp_fwd_kernel",
                 "mlp_bwd_rgb": "mlp_bwd_stashed_kernel", "mlp_bwd_alpha": "mlp_bwd_stashed_kernel",
                 "encode_rgb": "permuto_fwd_kernel", "encode_alpha": "permuto_fwd_kernel",
                 "lattice_bwd_rgb": "permuto_bwd_kernel", "lattice_bwd_alpha": "permuto_bwd_kernel"}
    per_kernel = {}
    for name, kern in kernel_of.items():
        per_kernel.setdefault(kern, []).append(name)
    dom_kernel = max(per_kernel, key=lambda k: sum(stages[n]["ms"] for n in per_kernel[k]))
    dom_stages = per_kernel[dom_kernel]
    dominant = dom_stages[0]
    dom_ms = statistics.mean(stages[n]["ms"] for n in dom_stages)               # average launch duration
    dom_alg = statistics.mean(stage_alg[n][1] for n in dom_stages)              # algorithmic flops / bytes of one launch
    dom_bound = stage_alg[dominant][0]
    if dom_kernel == "shells_trace_kernel":
        dom_bound = "hbm"  # BVH traversal is latency / L2 bound; against the HBM roofline its compulsory bytes are a lower bound only
    dom_peak, dom_unit, dom_div = (pk["hbm_gbs"], "GB/s", 1e9) if dom_bound == "hbm" else (pk["tflops_sustained"], "TFLOP/s", 1e12)
    dom_ach = dom_alg / (dom_ms * 1e-3) / dom_div
    roofline = {"bound": dom_bound, "achieved": round(dom_ach, 2), "peak": dom_peak, "unit": dom_unit, "frac": round(dom_ach / dom_peak, 4),
                "launches_per_step": len(dom_stages), "ms_per_launch": round(dom_ms, 4),
                "share_of_step": round(sum(stages[n]["ms"] for n in dom_stages) / sum(stage_ms), 3)}
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the stage's kernel from the ncu --set full captures of
    # this same workload at HEAD (profiles/r02_*.md); None where no capture exists
    ncu_traffic = NCU_TRAFFIC
    if dom_kernel in ("mlp_fwd_kernel", "mlp_bwd_stashed_kernel"):
        # the training-mode head kernels also stream the activation stash (what autograd keeps for the backward): their HBM view.
        # forward: features in + stash out + outputs; backward: stash in + upstream gradient + forward output in, feature gradient out
        stash_b = renderer.rgb_head.stash_bytes(n_hits)
        io = n_hits * POS_DIM * 4 + n_hits * 2 * 4 * 2  # features (or their gradient) + rgb/alpha outputs or gradients (3 + 1 floats, mean 2)
        hb = (stash_b + io) / (dom_ms * 1e-3) / 1e9
        roofline["hbm_view"] = {"bound": "hbm", "achieved": round(hb, 1), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(hb / pk["hbm_gbs"], 4),
                                "bytes_per_launch": int(stash_b + io),
                                "note": "same launch against the HBM roofline: activation stash (832 B per hit) + features + outputs"}
    roofline.update(kernel=dom_kernel, traffic=ncu_traffic.get(dom_kernel), peak_source=pk["source"],
                    note="algorithmic flops (or bytes) of one launch over the stage's mean CUDA-event duration inside the timed region "
                         "(event-record nodes of the replayed graph); traffic = DRAM bytes of one launch from the ncu capture in profiles/")

    # ---- the headline compositing kernels at the size SURVEY 8d prescribes (2^24 rays x 5, traffic >> L2)
    comp = None
    if not args.skip_composite_roofline:
        feats.clear(), dfeat.clear(), stash.clear()
        del loop, graph
        torch.cuda.empty_cache()
        n_big = 1 << 24
        d = all_hit_packed(n_big, K_LAYERS)
        rsp_b = RaySamplesPacked(0, 0, 0, 1)
        rsp_b.ray_start_end_idx = d["se"].to(dev)
        a, c, z = d["alpha"].to(dev), d["rgb"].to(dev), d["z"].to(dev)
        gs = [d[k].to(dev) for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")]
        for _ in range(3):
            VR.composite(rsp_b, a, c, z)
            VR.composite_backward(rsp_b, a, c, z, *gs)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        c0.record()
        for _ in range(reps):
            VR.composite(rsp_b, a, c, z)
            VR.composite_backward(rsp_b, a, c, z, *gs)
        c1.record()
        torch.cuda.synchronize()
        ms_c = c0.elapsed_time(c1) / reps
        nbytes = composite_bytes(n_big, n_big * K_LAYERS)
        comp = {"workload": "fused compositing fwd+bwd, 2^24 rays x 5 samples (5.8 GB algorithmic traffic per pair of launches)",
                "mrays_s": round(n_big / (ms_c * 1e-3) / 1e6, 1), "bound": "hbm", "achieved": round(nbytes / (ms_c * 1e-3) / 1e9, 1),
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(nbytes / (ms_c * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "ms_per_fwd_bwd": round(ms_c, 4)}

        # the long-ray case (BASELINE config[2]: NeRF-style packets, <= 1024 samples per ray): ring / scan kernel families
        del a, c, z, gs, d, rsp_b
        torch.cuda.empty_cache()
        p3 = nerf_packets(640000, seed_offset=3)
        rsp_l = RaySamplesPacked(0, 0, 0, 1)
        rsp_l.ray_start_end_idx = p3["se"].to(dev)
        a, c, z = p3["alpha"].to(dev), p3["rgb"].to(dev), p3["z"].to(dev)
        gs = [p3[k].to(dev) for k in ("g_rgb", "g_depth", "g_acc", "g_bgT")]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for _ in range(3):
            VR.composite(rsp_l, a, c, z)
            VR.composite_backward(rsp_l, a, c, z, *gs)
        ts = []
        for _ in range(7):
            flush.zero_()  # 1.2 GB of inputs per pass, but keep L2 cold between the timed pairs all the same
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            VR.composite(rsp_l, a, c, z)
            VR.composite_backward(rsp_l, a, c, z, *gs)
            c1.record()
            torch.cuda.synchronize()
            ts.append(c0.elapsed_time(c1))
        ms_l = sorted(ts)[len(ts) // 2]
        S_l = int(a.shape[0])
        nbytes_l = composite_bytes(640000, S_l)
        comp["long_rays"] = {"workload": f"fused compositing fwd+bwd, BASELINE config[2]: 640000 rays, {S_l} samples (<= 1024 per ray, 35 % empty rays)",
                             "mrays_s": round(640000 / (ms_l * 1e-3) / 1e6, 1), "bound": "hbm", "achieved": round(nbytes_l / (ms_l * 1e-3) / 1e9, 1),
                             "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(nbytes_l / (ms_l * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                             "ms_per_fwd_bwd": round(ms_l, 4)}
        del a, c, z, gs, p3, rsp_l, flush

    # ---- cpu baseline (bounded sample, rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.skip_cpu_baseline:
        path = CpuReferencePath(CPU_SAMPLE_RAYS)
        path.step()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            path.step()
        dt = time.perf_counter() - t0
        cpu = {"value": round(path.n_rays * reps / dt / 1e6, 4), "unit": UNIT, "cores": path.cores, "kind": "port",
               "sample": f"{reps} steps x {path.n_rays} rays (rows through the image centre) of the same workload; oracle port of the "
                         "reference algorithm (C BVH tracer with OpenMP, C permutohedral encoders with OpenMP, torch CPU heads, dense torch compositing, autograd "
                         "through compositing and heads, encoder backward)"}
        cpu["compositing_c1"] = cpu_compositing_c1()

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (ray/triangle, packing, compositing); f16 operands x f32 accumulate (MLP heads)", "data": "synthetic",
        "config": workload_config(args, world) | {"hits_per_step_all_ranks": n_hits_all, "allreduce_bytes_per_step": allreduce_bytes if world > 1 else 0,
                                                  "allreduce_in_graph": bool(reduce_in_graph) if world > 1 else None},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(N * 24), "d2h_bytes_per_step": int(N * 12 + 4) if args.e2e_image else 4,
                "ms_per_step": round(ms_e2e / args.steps, 4), "mode": e2e_mode,
                "note": "pinned host rays -> device every step, " + ("composited image + loss" if args.e2e_image else "the step's loss")
                        + " -> pinned host every step (--e2e-image adds the composited image, 12 B per ray); everything between (hits, encoder "
                          "features, head outputs, gradients) is produced and consumed on the device"},
        "gpu_launches": int(launches),
        "launch_mode": ("one CUDA graph per step (%d kernels of this library per step + torch elementwise ops)" % kernels_per_step) if use_graph
                       else "eager (one Python call per kernel)",
        "roofline": roofline,
        "roofline_hbm": None if comp is None else {
            "kernel": "composite_fwd_tile_kernel<TMA> + composite_bwd_tile_kernel<TMA>", "bound": "hbm", "achieved": comp["achieved"],
            "peak": comp["peak"], "unit": "GB/s", "frac": comp["frac"], "traffic": NCU_TRAFFIC.get("composite_tile_fwd+bwd"),
            "note": "BASELINE's '% of HBM roofline': fused compositing fwd+bwd at 2^24 rays x 5 samples, algorithmic bytes 64 + 56 s per ray "
                    "(SURVEY 8d) over CUDA-event time; traffic = DRAM bytes of the pair of launches from the ncu capture"},
        "stages": stages,
        "compositing_roofline": comp,
        "cpu_baseline": cpu,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-composite-roofline", action="store_true")
    ap.add_argument("--e2e-image", action="store_true", help="e2e also copies the composited image back to the host every step (default: "
                    "the step's result is its loss)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if world > 1:
        # NCCL prints its version banner on STDOUT when the box exports NCCL_DEBUG=VERSION; stdout carries the one JSON line only
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
