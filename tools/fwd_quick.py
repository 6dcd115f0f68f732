import os, sys, torch
sys.path.insert(0, os.getcwd())
from mdqe_cvpr2023_b200 import ops
from tests.gpu_util import R50_360, R50_720, make_inputs, to_cuda
flush = torch.ones(160 * 1024 * 1024, device="cuda")
def timed(fn, iters=15):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.sum(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
out = []
for sname, pyr in (("R50_360", R50_360), ("R50_720", R50_720)):
    for dist in ("local", "uniform"):
        inp = to_cuda(make_inputs(4, pyr, 8, 32, 4, dist=dist, seed=0))
        a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
        out.append(f"{sname}/{dist} {timed(lambda: ops.ms_deform_attn_forward(*a, 64)):6.1f}")
print(os.environ.get("MSDA_B200_LIB", "default")[-12:], " | ".join(out), flush=True)
