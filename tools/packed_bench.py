#!/usr/bin/env python
"""bf16 forward: reference layout (msda_forward, MSDA_BF16 / BF16_LOC32) vs the paired-corner layout (msda_pack_value +
msda_forward_packed) at the encoder shapes.  CUDA events around each call, L2 flushed between launches, median."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mdqe_cvpr2023_b200 import ops  # noqa: E402
from tests.gpu_util import R50_360, R50_720, make_inputs, to_cuda  # noqa: E402

flush = torch.ones(160 * 1024 * 1024, device="cuda")


def timed(fn, iters=15):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for sname, pyr, Lq in (("R50_360 enc", R50_360, None), ("R50_720 enc", R50_720, None), ("R50_360 dec (196 q)", R50_360, 196)):
    for dist in ("local", "uniform"):
        inp = to_cuda(make_inputs(4, pyr, 8, 32, 4, Lq=Lq, dist=dist, seed=0))
        v32, vbf = inp["value"], inp["value"].bfloat16()
        for ln, ldt in (("loc bf16", torch.bfloat16), ("loc fp32", torch.float32)):
            loc, aw = inp["loc"].to(ldt), inp["aw"].to(ldt)
            base = timed(lambda: ops.ms_deform_attn_forward(vbf, inp["shapes"], inp["level_start"], loc, aw, 64))
            f32 = timed(lambda: ops.ms_deform_attn_forward(v32, inp["shapes"], inp["level_start"], inp["loc"], inp["aw"], 64))
            pk_bf = timed(lambda: ops.pack_value(vbf, inp["shapes"], inp["level_start"]))
            pk_32 = timed(lambda: ops.pack_value(v32, inp["shapes"], inp["level_start"]))
            packed = ops.pack_value(vbf, inp["shapes"], inp["level_start"])
            smp = timed(lambda: ops.ms_deform_attn_forward_packed(packed, vbf.shape, inp["shapes"], inp["level_start"], loc, aw))
            res[f"{sname}/{dist}/{ln}"] = dict(fp32_us=f32, bf16_reference_layout_us=base, pack_from_bf16_us=pk_bf, pack_from_fp32_us=pk_32, packed_forward_us=smp)
            print(f"{sname:20s} {dist:8s} {ln}: fp32 {f32:6.1f} | bf16 reference layout {base:6.1f} | pack (bf16 in) {pk_bf:5.1f} (fp32 in) {pk_32:5.1f} | packed forward {smp:6.1f}", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "packed_bench.json"), "w"), indent=1)
