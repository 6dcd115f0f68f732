#!/usr/bin/env python
"""Host<->device copy bandwidth with N ranks copying AT THE SAME TIME (what bounds the e2e arm of bench.py at N > 1).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/hostlink_probe.py

Every rank: 512 MB pinned buffer, H2D alone, D2H alone, both directions at once (two streams); all ranks start behind a barrier.
Rank 0 prints per-rank and aggregate GB/s plus the NUMA node of every GPU and the CPUs this process may run on."""
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 512 << 20
h1, h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
h1.fill_(1); h2.fill_(2)
d1, d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev), torch.empty(nbytes, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def run(kind, reps=6):
    def once():
        if kind in ("h2d", "duplex"):
            d1.copy_(h1, non_blocking=True)
        if kind in ("d2h", "duplex"):
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    once(); sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return (2 if kind == "duplex" else 1) * reps * nbytes / dt / 1e9


p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    numa = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
except OSError:
    numa = "?"
res = [run(k) for k in ("h2d", "d2h", "duplex")]
row = torch.tensor(res + [float(numa) if numa.lstrip("-").isdigit() else -1.0, float(len(os.sched_getaffinity(0)))], device=dev, dtype=torch.float64)
rows = [torch.zeros_like(row) for _ in range(world)]
if world > 1:
    dist.all_gather(rows, row)
else:
    rows = [row]
if rank == 0:
    print(f"{world} rank(s) copying concurrently, 512 MB pinned buffers, GB/s per rank:")
    print("rank  numa  cpus   h2d    d2h   duplex(sum)")
    for r, t in enumerate(rows):
        t = t.tolist()
        print(f"{r:4d}  {int(t[3]):4d}  {int(t[4]):4d}  {t[0]:5.1f}  {t[1]:5.1f}  {t[2]:6.1f}")
    agg = torch.stack(rows).sum(0).tolist()
    print(f"aggregate         {agg[0]:6.1f} {agg[1]:6.1f} {agg[2]:7.1f}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
