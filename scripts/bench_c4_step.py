"""BASELINE config[3]: the 5-mesh volsurfs TRAINING step with the legacy appearance — permutohedral hash encoding (24 levels x 2, 2^18
entries, one encoder per head as in volsurfs_py/models/rgb.py:40-60) + 64-wide heads — on 2^18 rays per step sharded over the ranks
(32768 rays per GPU at 8 GPUs), followed by the NCCL all-reduce of the head AND lattice gradients (2 x 50.3 MB per step).

    python scripts/bench_c4_step.py [--rays-per-gpu 32768] [--steps 20]                                   (one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_c4_step.py   (N GPUs)

Stages through the library's autograd wrappers: ShellTracer.render_samples -> PermutoHashEncoder x 2 -> AppearanceHead x 2 ->
CompositeFunc -> L1 loss -> backward -> GradAllReducer.  Timed with CUDA events between barriers, max over ranks; rank 0 prints one JSON
line (Mrays/s over all ranks; the all-reduce share is measured by running the same steps with the exchange switched off)."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from volsurfs_b200.appearance import AppearanceHead  # noqa: E402
from volsurfs_b200.dist import GradAllReducer  # noqa: E402
from volsurfs_b200.encoding import PermutoHashEncoder  # noqa: E402
from volsurfs_b200.raytracer import ShellTracer  # noqa: E402
from volsurfs_b200.synthetic import camera_rays, shell_meshes  # noqa: E402
from volsurfs_b200.volume_rendering import composite  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays-per-gpu", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="explicit-kernel step replayed as one CUDA graph")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tracer = ShellTracer(shell_meshes(K=5))
    torch.manual_seed(7)  # same initial parameters on every rank
    encs = [PermutoHashEncoder(bb_sides=2.0, device=dev) for _ in range(2)]
    heads = [AppearanceHead(encs[0].output_dim, (64, 64, 64), 3, 3, False, "gelu", False).to(dev),
             AppearanceHead(encs[1].output_dim, (64, 64, 64), 1, 3, False, "gelu", True).to(dev)]
    params = [p for e in encs for p in e.parameters()] + [p for h in heads for p in h.parameters()]
    grad_bytes = sum(p.numel() * 4 for p in params)
    o_all, d_all = camera_rays(800, 800, azimuth_deg=30.0 + 40.0 * rank)
    g = torch.Generator().manual_seed(100 + rank)
    reducer = GradAllReducer()

    def batch():
        idx = torch.randint(0, o_all.shape[0], (args.rays_per_gpu,), generator=g)
        return o_all[idx].to(dev), d_all[idx].to(dev), torch.rand(args.rays_per_gpu, 3, generator=g).to(dev)

    def step(o, d, gt, exchange=True):
        for p in params:
            p.grad = None
        rsp = tracer.render_samples(o, d, exact_size=True, with_normals=True)
        f_rgb, _ = encs[0](rsp.samples_3d)
        f_alpha, _ = encs[1](rsp.samples_3d)
        rgb = heads[0](f_rgb, rsp.samples_dirs, rsp.samples_normals)
        alpha = heads[1](f_alpha, rsp.samples_dirs, rsp.samples_normals)
        rgb_fg, _, _, bgT = composite(rsp, alpha, rgb)
        loss = (rgb_fg + bgT - gt).abs().mean()
        loss.backward()
        if exchange and world > 1:
            reducer.launch([p.grad for p in params])
            reducer.wait()
        return loss, rsp.get_total_nr_samples()

    batches = [batch() for _ in range(4)]
    if args.graph:
        # the same step through EncodedShellRenderer (explicit forward / backward kernels, capacity-sized buffers, no host read),
        # captured once into a CUDA graph; the gradient exchange runs on the replayed graph's outputs
        from volsurfs_b200.pipeline import EncodedShellRenderer

        er = EncodedShellRenderer(tracer, heads[0], heads[1], encs[0], encs[1])
        s_o, s_d, s_gt = (t.clone() for t in batches[0])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                er.render_fwd_bwd(s_o, s_d, s_gt)
        torch.cuda.current_stream().wait_stream(side)
        # two graphs: everything up to the colour branch's gradients, then the transparency branch's backward — the all-reduce of the
        # first 50 MB lattice gradient runs on NCCL's stream under the second graph
        pool = torch.cuda.graph_pool_handle()
        graph1, graph2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph1, pool=pool):
            g_out = er.render_fwd_bwd(s_o, s_d, s_gt, rgb_branch_only=True)
        with torch.cuda.graph(graph2, pool=pool):
            g_out.update(er.branch_backward("alpha"))

        def step(o, d, gt, exchange=True):  # noqa: F811
            s_o.copy_(o)
            s_d.copy_(d)
            s_gt.copy_(gt)
            graph1.replay()
            if exchange and world > 1:
                reducer.launch([g_out["grad_lattice_rgb"], g_out["grad_rgb"]])
            graph2.replay()
            if exchange and world > 1:
                reducer.launch([g_out["grad_lattice_alpha"], g_out["grad_alpha"]])
                reducer.wait()
            return g_out["loss"], 0

    for i in range(args.warmup):
        step(*batches[i % 4])

    def timed(exchange):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hits = 0
        for i in range(args.steps):
            _, s = step(*batches[i % 4], exchange=exchange)
            hits += s
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]) / args.steps, hits / args.steps

    ms_step, hits = timed(True)
    ms_noex, _ = timed(False)
    if rank == 0:
        print(json.dumps({
            "workload": "BASELINE config[3]: 5-mesh volsurfs training step, permutohedral encoding (24x2, 2^18) + 64-wide heads",
            "n_gpus": world, "rays_per_gpu": args.rays_per_gpu, "rays_per_step": args.rays_per_gpu * world, "hits_per_gpu": round(hits),
            "ms_per_step": round(ms_step, 3), "mrays_s": round(args.rays_per_gpu * world / ms_step / 1e3, 2),
            "ms_per_step_without_exchange": round(ms_noex, 3), "allreduce_bytes_per_step": grad_bytes if world > 1 else 0,
            "mode": "EncodedShellRenderer step replayed as one CUDA graph" if args.graph else "autograd-driven (eager)",
            "note": "eager mode reads the sample count on the host once per step (exact-size packing); all-reduce = fp32 sum of every lattice "
                    "and head gradient, 64 MB buckets, mean over ranks"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
