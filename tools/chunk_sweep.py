import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import _lib, ops
from tests.gpu_util import R50_360, make_inputs, to_cuda
from tools.kernel_bench import timed
flush = torch.empty(128 * 1024 * 1024, device="cuda")
inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, dist="local", seed=0))
a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
for chunk in (16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256):
    _lib.set_option("chunk_pairs", chunk)
    f = timed(lambda: ops.ms_deform_attn_forward(*a, 64), 15, flush)
    b = timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), 15, flush)
    print(f"chunk {chunk:4d}: fwd {f['median_us']:.1f} us  bwd(+memset) {b['median_us']:.1f} us", flush=True)
