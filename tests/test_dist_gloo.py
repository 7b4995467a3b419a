"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2 (ray sharding, gradient all-reduce semantics,
row gather).  The GPU box runs the same code over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from volsurfs_b200.dist import GradAllReducer, allreduce_gradients, gather_rows, shard_range, shard_rays


def test_shard_range_partitions():
    for n in (0, 1, 7, 640000, 262144, 1920000):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rays):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays_o = torch.randn(n_rays, 3, generator=g)
        rays_d = torch.randn(n_rays, 3, generator=g)
        feats = torch.randn(n_rays, 8, generator=g)
        target = torch.randn(n_rays, 3, generator=g)
        torch.manual_seed(1)
        net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 3))
        # single-process gradient on the full batch (loss = mean over the local batch, utils/losses.py:18)
        full = torch.nn.functional.l1_loss(net(feats), target)
        want = torch.autograd.grad(full, list(net.parameters()))
        # sharded: equal blocks, local mean loss, mean all-reduce
        s, c = shard_range(n_rays, rank, world)
        o_loc, d_loc = shard_rays(rays_o, rays_d, rank, world)
        assert o_loc.shape[0] == c and torch.equal(o_loc, rays_o[s:s + c]) and torch.equal(d_loc, rays_d[s:s + c])
        loss = torch.nn.functional.l1_loss(net(feats[s:s + c]), target[s:s + c])
        loss.backward()
        allreduce_gradients(net.parameters(), bucket_bytes=256, local_count=c)  # tiny buckets: exercises the bucketing
        for p, w in zip(net.parameters(), want):
            assert torch.allclose(p.grad, w, atol=1e-6), (p.grad - w).abs().max()
        # explicit launch/wait with several groups
        r = GradAllReducer(bucket_bytes=64)
        a = torch.full((5,), float(rank + 1))
        b = torch.full((3, 2), float(10 * (rank + 1)))
        r.launch([a])
        r.launch([b, None])
        r.wait()
        assert torch.allclose(a, torch.full((5,), 1.5)) and torch.allclose(b, torch.full((3, 2), 15.0))
        # row gather of rank-local images
        img = torch.arange(n_rays * 3, dtype=torch.float32).view(n_rays, 3)
        got = gather_rows(img[s:s + c].clone(), n_rays)
        assert torch.equal(got, img)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [64, 101])
def test_gloo_world2(n_rays):
    mp.spawn(_worker, args=(2, _free_port(), n_rays), nprocs=2, join=True)
