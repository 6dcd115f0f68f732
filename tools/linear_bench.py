"""Tensor-core Linear (3xTF32) vs torch F.linear (cuBLAS fp32, TF32 off as in the reference) at the module's shapes.
CUDA events, L2 flushed between launches."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from mdqe_cvpr2023_b200 import _lib, ops

if "--tiles" in sys.argv:                        # A/B: whole tiles dealt round-robin instead of contiguous (tile, chunk) ranges
    _lib.set_option("gemm_stream_k", 0)

torch.backends.cuda.matmul.allow_tf32 = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def timed_graph(fn, reps=20):
    """GPU time without host dispatch: 10 calls captured in one CUDA graph (L2 warm, as inside a training step)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (10 * reps)


res = []
for name, rows, in_f, out_f in [("enc value/output/offsets proj (R50_360, T=4)", 4 * 5100, 256, 256), ("enc attention_weights", 4 * 5100, 256, 128),
                                ("R50_720 value proj", 4 * 15300, 256, 256), ("swinl value proj (T=3)", 3 * 5100, 192, 192),
                                ("decoder query proj (196 x T)", 784, 256, 256)]:
    x = torch.randn(rows, in_f, device="cuda"); w = torch.randn(out_f, in_f, device="cuda") / 16; b = torch.randn(out_f, device="cuda")
    gy = torch.randn(rows, out_f, device="cuda")
    r = {"shape": name, "rows": rows, "in": in_f, "out": out_f,
         "tc_fwd_us": timed(lambda: ops.tc_linear_forward(x, w, b)), "torch_fwd_us": timed(lambda: F.linear(x, w, b)),
         "tc_dgrad_us": timed(lambda: ops.tc_linear_backward(gy, x, w, True, False)), "torch_dgrad_us": timed(lambda: gy @ w),
         "tc_wgrad_us": timed(lambda: ops.tc_linear_backward(gy, x, w, False, True)), "torch_wgrad_us": timed(lambda: gy.t() @ x),
         "tc_wgrad_bias_us": timed(lambda: ops.tc_linear_backward(gy, x, w, False, True, True))}
    r.update({"tc_fwd_graph_us": timed_graph(lambda: ops.tc_linear_forward(x, w, b)), "torch_fwd_graph_us": timed_graph(lambda: F.linear(x, w, b)),
              "tc_dgrad_graph_us": timed_graph(lambda: ops.tc_linear_backward(gy, x, w, True, False)), "torch_dgrad_graph_us": timed_graph(lambda: gy @ w),
              "tc_wgrad_graph_us": timed_graph(lambda: ops.tc_linear_backward(gy, x, w, False, True)), "torch_wgrad_graph_us": timed_graph(lambda: gy.t() @ x),
              "tc_wgrad_bias_graph_us": timed_graph(lambda: ops.tc_linear_backward(gy, x, w, False, True, True))})
    flops = 2.0 * rows * in_f * out_f
    r["tc_fwd_tflops"] = flops / r["tc_fwd_us"] / 1e6
    r["fwd_bytes_GBps"] = (rows * (in_f + out_f) * 4) / r["tc_fwd_us"] / 1e3
    res.append(r)
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/linear_bench%s.json" % ("_tiles" if "--tiles" in sys.argv else ""), "w"), indent=1)
