#!/usr/bin/env python
"""Gradient all-reduce in isolation: NCCL vs msda_allreduce_f32 (two-shot P2P / multimem), CUDA events, max over ranks.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/allreduce_bench.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdqe_cvpr2023_b200.collectives import PeerAllReduce  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
N = 19_480_000                                              # enc + dec parameters of MDQE: 77.9 MB fp32


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


x = torch.zeros(N, device=dev)
ms = timed(lambda: dist.all_reduce(x))
if rank == 0:
    print(f"world {world}: NCCL all_reduce of {N * 4 / 1e6:.1f} MB: {ms * 1e3:8.1f} us  ({N * 4 / ms / 1e6:7.1f} GB/s algorithmic)", flush=True)
ar = PeerAllReduce(N, dev, algo="p2p")
for algo in ["p2p"] + (["multimem"] if ar._mc else []):
    ar.algo = algo
    for ctas in (1, 2, 4, 8, 16, 32):
        ms = timed(lambda: ar.all_reduce_(mean=True, n_ctas=ctas))
        ar.check()
        if rank == 0:
            print(f"world {world}: {algo:9s} {ctas:3d} CTAs: {ms * 1e3:8.1f} us  ({N * 4 / ms / 1e6:7.1f} GB/s algorithmic)", flush=True)
dist.barrier()
dist.destroy_process_group()
