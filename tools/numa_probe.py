"""Probe: does the NUMA node of a pinned host buffer change H2D / D2H bandwidth on this box?
Prints the topology, then GB/s per node (affinity set before the pinned allocation = first touch on that node)."""
import glob, os, re, subprocess, time
import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    nodes[int(re.findall(r"\d+$", d)[0])] = cpulist(open(d + "/cpulist").read())
print("nodes:", {k: (len(v), v[:2]) for k, v in nodes.items()})
print("affinity now:", len(os.sched_getaffinity(0)), "cpus; cpu_count", os.cpu_count())
p = torch.cuda.get_device_properties(0)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    print("gpu", bus, "numa_node", open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
except Exception as e:
    print("numa_node unreadable:", e)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
full = os.sched_getaffinity(0)
nbytes = 512 << 20
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for node, cpus in nodes.items():
    cpus = [c for c in cpus if c in full]
    if not cpus:
        print("node", node, "no allowed cpus")
        continue
    os.sched_setaffinity(0, cpus)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    for name, fn in (("h2d", lambda: dev.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(dev, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        print("node", node, name, "%.1f GB/s" % (5 * nbytes / (time.perf_counter() - t0) / 1e9))
    s2 = torch.cuda.Stream()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dev.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(dev2, non_blocking=True)
    torch.cuda.synchronize()
    print("node", node, "duplex", "%.1f GB/s (sum)" % (10 * nbytes / (time.perf_counter() - t0) / 1e9))
    del h, h2
    os.sched_setaffinity(0, full)
