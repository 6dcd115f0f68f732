"""Clip sharding across the GPUs of one box.

Every batch element of the operator is independent (frames in the encoder, clips in the decoder and
mask head), so multi-GPU execution is one process per GPU, each owning a disjoint set of clips, weights
replicated.  Inference needs no communication; training adds one gradient all-reduce per step (PyTorch
DDP over NCCL/NVLink, as detectron2's launch does in the reference, train_net.py:256-271; evaluation
shards videos with InferenceSampler, mdqe/data/build.py:245).
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [begin, end) of `n_items` for `rank` (first n_items % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_round_robin(n_items, rank, world):
    """Indices rank, rank+world, ... (how InferenceSampler-style video sharding interleaves long and short videos)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gather_clip_outputs(local, n_items, group=None):
    """All-gather per-clip outputs sharded with shard_bounds back into clip order on every rank.
    `local` is [n_local, ...]; shards may differ in length by one."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    longest = (n_items + world - 1) // world
    pad = local.new_zeros((longest,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r in range(world):
        b, e = shard_bounds(n_items, r, world)
        out.append(parts[r][: e - b])
    return torch.cat(out)


def allreduce_mean_gradients(parameters, group=None, bucket_bytes=64 << 20):
    """Average .grad over the ranks in flat buckets (what DDP does; for loops that do not wrap the model in
    DistributedDataParallel).  Buckets are sized for launch latency, not for link count: NVSwitch gives
    every GPU full bandwidth to every peer.

    Every rank must issue the same collectives on the same sizes, so the buckets are built from ALL parameters that require
    a gradient, in the order given, whether or not this rank produced one: a parameter unused on this rank (e.g.
    `sampling_grid_offsets` when a clip has no matched instance) contributes zeros and receives the mean of the other ranks'
    gradients as its `.grad`.  Buckets never mix dtypes (torch.cat would promote and the copy back would truncate)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    params = [p for p in parameters if p.requires_grad]
    n_buckets, i = 0, 0
    while i < len(params):
        bucket, size, dtype = [], 0, params[i].dtype
        while i < len(params) and params[i].dtype == dtype and \
                (not bucket or size + params[i].numel() * params[i].element_size() <= bucket_bytes):
            bucket.append(params[i])
            size += params[i].numel() * params[i].element_size()
            i += 1
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, group=group)
        flat.div_(world)
        off = 0
        for p in bucket:
            piece = flat[off: off + p.numel()].view_as(p)
            if p.grad is None:
                p.grad = piece.clone()
            else:
                p.grad.copy_(piece)
            off += p.numel()
        n_buckets += 1
    return n_buckets
