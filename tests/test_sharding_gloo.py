"""Multi-process host logic on CPU (gloo, world_size 2): clip sharding covers every clip once, gathered
per-shard outputs equal the single-process result, and data-parallel gradients (DDP and the bucketed
all-reduce helper) equal the single-process gradients on the whole batch.  The module runs on the
C-oracle Function (no GPU here); the N>1 CUDA path is the same code with the CUDA Function."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mdqe_cvpr2023_b200.sharding import shard_bounds, shard_round_robin


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 9, 30):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = shard_bounds(n, r, world)
                assert 0 <= b <= e <= n and (e - b) in (n // world, n // world + 1)
                seen += list(range(b, e))
            assert seen == list(range(n))
            rr = sorted(i for r in range(world) for i in shard_round_robin(n, r, world))
            assert rr == list(range(n))
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_problem():
    import mdqe_cvpr2023_b200.modules as M
    from tests.helpers import OracleMSDAFunction
    M.MSDeformAttnFunction = OracleMSDAFunction
    torch.manual_seed(0)
    mod = M.MSDeformAttn(d_model=32, n_levels=2, n_heads=4, n_points=2, pred_offsets=True, mode="spatial")
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in mod.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))
    shapes = torch.tensor([(4, 6), (2, 3)])
    S = 30
    clips = 5                                               # odd on purpose: shards of 3 and 2
    x = torch.randn(clips, S, 32, generator=g)
    ref = torch.cat([torch.rand(clips, S, 2, generator=g), torch.full((clips, S, 2), 0.1)], -1)
    tgt = torch.randn(clips, S, 32, generator=g)
    return mod, x, ref, shapes, tgt


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdqe_cvpr2023_b200.sharding import allreduce_mean_gradients, gather_clip_outputs
        mod, x, ref, shapes, tgt = _make_problem()
        b, e = shard_bounds(x.shape[0], rank, world)
        # (a) inference: no communication until the gather
        with torch.no_grad():
            local = mod(x[b:e], ref[b:e], x[b:e], shapes)
        full = gather_clip_outputs(local, x.shape[0])
        # (b) training with the bucketed all-reduce helper; per-clip losses summed, then averaged over ranks
        mod.zero_grad()
        ((mod(x[b:e], ref[b:e], x[b:e], shapes) - tgt[b:e]) ** 2).sum().backward()
        nb = allreduce_mean_gradients(list(mod.parameters()), bucket_bytes=4096)
        helper = {k: p.grad.clone() for k, p in mod.named_parameters()}
        # (b') a parameter that only rank 0 used (and of another dtype): same collectives on every rank, no hang, and the
        # ranks that had no gradient receive the mean
        extra = torch.nn.Linear(4, 3).double()
        with torch.no_grad():
            for j, p in enumerate(extra.parameters()):
                p.copy_(torch.arange(p.numel(), dtype=torch.float64).view_as(p) * 0.1 + j)
        xin = torch.arange(8.0, dtype=torch.float64).view(2, 4)
        g0 = torch.autograd.grad((extra(xin) ** 2).sum(), list(extra.parameters()))
        if rank == 0:
            (extra(xin) ** 2).sum().backward()
        mod.zero_grad()
        ((mod(x[b:e], ref[b:e], x[b:e], shapes) - tgt[b:e]) ** 2).sum().backward()
        allreduce_mean_gradients(list(extra.parameters()) + list(mod.parameters()), bucket_bytes=4096)
        for p_, g_ in zip(extra.parameters(), g0):
            assert p_.grad is not None and p_.grad.dtype == torch.float64 and torch.allclose(p_.grad, g_ / world), "unused-parameter bucket"
        for k, p_ in mod.named_parameters():
            assert torch.allclose(p_.grad, helper[k], atol=1e-6), k
        # (c) training under DistributedDataParallel (equal shard sizes needed: use the first 4 clips)
        b4, e4 = shard_bounds(4, rank, world)
        ddp = torch.nn.parallel.DistributedDataParallel(mod)
        ddp.zero_grad()
        ((ddp(x[b4:e4], ref[b4:e4], x[b4:e4], shapes) - tgt[b4:e4]) ** 2).sum().backward()
        ddp_grads = {k: p.grad.clone() for k, p in mod.named_parameters()}
        # (d) the same under DDP with collectives.make_ddp_comm_hook: the hook's bucket plumbing (copy in, reduce a multiple of 4
        # floats, copy back, completed future) on a stand-in for PeerAllReduce whose "kernel" is gloo's all_reduce -- the real
        # kernel needs NVLink peers (tests/test_ddp_nccl_gpu.py[ddp_hook], tests/test_allreduce_gpu.py)
        from mdqe_cvpr2023_b200.collectives import make_ddp_comm_hook

        class GlooStandIn:
            def __init__(self, numel):
                self.numel = numel
                self.buffer = torch.zeros(numel)
                self.calls = 0

            def all_reduce_(self, offset, numel, mean=True):
                assert offset % 4 == 0 and numel % 4 == 0, "msda_allreduce_f32 takes multiples of 4 floats"
                seg = self.buffer[offset:offset + numel]
                dist.all_reduce(seg)
                if mean:
                    seg.div_(world)
                self.calls += 1

        stand_in = GlooStandIn(sum(p.numel() for p in mod.parameters()) + 8)
        ddp2 = torch.nn.parallel.DistributedDataParallel(mod)
        ddp2.register_comm_hook(None, make_ddp_comm_hook(stand_in))
        ddp2.zero_grad()
        ((ddp2(x[b4:e4], ref[b4:e4], x[b4:e4], shapes) - tgt[b4:e4]) ** 2).sum().backward()
        assert stand_in.calls >= 1, "the hook did not run"
        for k, p_ in mod.named_parameters():
            assert torch.allclose(p_.grad, ddp_grads[k], atol=1e-6), f"comm hook: {k}"
        if rank == 0:
            # plain numpy through the queue (tensor fd-sharing does not survive the worker's exit)
            q.put(dict(full=full.numpy(), helper={k: v.numpy() for k, v in helper.items()},
                       ddp={k: v.numpy() for k, v in ddp_grads.items()}, buckets=nb))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    mod, x, ref, shapes, tgt = _make_problem()
    with torch.no_grad():
        want = mod(x, ref, x, shapes)
    got = dict(full=torch.from_numpy(got["full"]), helper={k: torch.from_numpy(v) for k, v in got["helper"].items()},
               ddp={k: torch.from_numpy(v) for k, v in got["ddp"].items()}, buckets=got["buckets"])
    assert torch.allclose(got["full"], want, atol=1e-6)
    mod.zero_grad()
    ((mod(x, ref, x, shapes) - tgt) ** 2).sum().backward()
    for k, p in mod.named_parameters():
        assert torch.allclose(got["helper"][k], p.grad / world, atol=1e-5, rtol=1e-4), k
    assert got["buckets"] >= 2
    mod.zero_grad()
    ((mod(x[:4], ref[:4], x[:4], shapes) - tgt[:4]) ** 2).sum().backward()
    for k, p in mod.named_parameters():
        assert torch.allclose(got["ddp"][k], p.grad / world, atol=1e-5, rtol=1e-4), k
