"""NCCL all-reduce (mean, fp32) at the gradient sizes of the training step: time per call and algorithm bandwidth.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_allreduce.py"""
import json
import os

import torch
import torch.distributed as dist

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = dist.get_world_size()
for mb in (0.135, 12.6, 25.2, 50.3, 100.7):
    t = torch.randn(int(mb * 1e6 / 4), device="cuda")
    for _ in range(5):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mbytes": mb, "ms": round(ms, 4), "algbw_gbs": round(mb / ms, 1),
                          "busbw_gbs": round(mb / ms * 2 * (world - 1) / world, 1), "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}),
              flush=True)
dist.destroy_process_group()
