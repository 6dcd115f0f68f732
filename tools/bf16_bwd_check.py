"""bf16 backward timing at the encoder shape (L2 flushed), next to fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import ops
from tests.gpu_util import R50_360, make_inputs, to_cuda
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=10):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3
for D, N in ((32, 4), (24, 3)):
    inp = to_cuda(make_inputs(N, R50_360, 8, D, 4, dist="local", seed=0))
    v, sh, ls, loc, aw, go = (inp[k] for k in ("value", "shapes", "level_start", "loc", "aw", "grad_out"))
    f32 = timed(lambda: ops.ms_deform_attn_backward(v, sh, ls, loc, aw, go, 64))
    b = [t.bfloat16() for t in (v, loc, aw, go)]
    b16 = timed(lambda: ops.ms_deform_attn_backward(b[0], sh, ls, b[1], b[2], b[3], 64))
    b16l = timed(lambda: ops.ms_deform_attn_backward(b[0], sh, ls, loc, aw, b[3], 64))
    print(f"D={D} N={N}: bwd fp32 {f32:.1f} us | bf16 {b16:.1f} us | bf16 values + fp32 loc/aw {b16l:.1f} us")
