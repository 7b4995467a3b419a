"""Multi-GPU plumbing for the hot path: rays are independent, so they are sharded in contiguous blocks over ranks (one process
per GPU, ``torch.distributed``; NCCL over NVLink on the GPU box, gloo in CPU tests).  Packed buffers, ``ray_start_end_idx`` offsets
and per-ray outputs are rank-local — rendering needs no data-path collective.  Training adds ONE exchange: the all-reduce of the
appearance-head (and encoding) gradients, bucketed and launched asynchronously so it overlaps the remaining backward work.

The reference has no distributed code at all; its loss is a mean over the local batch
(volsurfs_py/utils/losses.py:18), so summed gradients are divided by the world size."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block [start, start+count) of rank ``rank``; blocks differ by at most one ray and cover [0, n)"""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def shard_rays(rays_o: torch.Tensor, rays_d: torch.Tensor, rank: int, world: int):
    s, c = shard_range(rays_o.shape[0], rank, world)
    return rays_o[s:s + c], rays_d[s:s + c]


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradAllReducer:
    """Bucketed asynchronous gradient all-reduce (mean).  ``launch(tensors)`` can be called as soon as a group of gradients is
    final (e.g. per head); ``wait()`` blocks the current stream until every bucket has been reduced and copied back."""

    def __init__(self, bucket_bytes: int = 64 << 20, group=None):
        self.bucket_bytes = bucket_bytes
        self.group = group
        self._pending: List[tuple] = []

    def launch(self, tensors: Iterable[torch.Tensor]) -> None:
        world = _world()
        tensors = [t for t in tensors if t is not None]
        if world == 1 or not tensors:
            return
        bucket, size = [], 0
        for t in tensors:
            nbytes = t.numel() * t.element_size()
            if nbytes * 2 >= self.bucket_bytes and t.is_contiguous():
                # a tensor of bucket size on its own (a hash-table gradient: 50 MB): reduced in place, no staging copy in or out
                work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._pending.append((work, t, None, world))
                continue
            if bucket and size + nbytes > self.bucket_bytes:
                self._flush(bucket, world)
                bucket, size = [], 0
            bucket.append(t)
            size += nbytes
        if bucket:
            self._flush(bucket, world)

    def _flush(self, bucket, world):
        flat = torch.cat([t.reshape(-1) for t in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, flat, bucket, world))

    def wait(self) -> None:
        for work, flat, bucket, world in self._pending:
            work.wait()
            flat.div_(world)
            if bucket is None:
                continue
            off = 0
            for t in bucket:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        self._pending.clear()


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, local_count=None) -> None:
    """mean-all-reduce ``p.grad`` of every parameter that has one.  With equal shards the plain mean reproduces the gradient of
    the global-batch mean loss; pass ``local_count`` (rays of this rank) when shards differ in size so that every rank's
    local-mean gradient is weighted by its share of the global batch."""
    grads = [p.grad for p in params if p.grad is not None]
    world = _world()
    if world > 1 and local_count is not None and grads:
        total = torch.tensor([float(local_count)], dtype=torch.float64, device=grads[0].device)
        dist.all_reduce(total)
        scale = float(local_count) * world / float(total.item())
        for g in grads:
            g.mul_(scale)
    r = GradAllReducer(bucket_bytes)
    r.launch(grads)
    r.wait()


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """full [n_total, d] tensor from the ranks' contiguous row blocks (full-frame render: 12 bytes per ray)"""
    world = _world()
    if world == 1:
        return local
    rank = dist.get_rank()
    counts = [shard_range(n_total, r, world)[1] for r in range(world)]
    assert local.shape[0] == counts[rank]
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
