import sys, torch
sys.path.insert(0, "/root/repo")
from mdqe_cvpr2023_b200 import ops
from tests.gpu_util import R50_360, make_inputs, to_cuda
inp = to_cuda(make_inputs(4, R50_360, 8, 32, 4, dist="local", seed=0))
vbf = inp["value"].bfloat16(); loc = inp["loc"].bfloat16(); aw = inp["aw"].bfloat16()
packed = ops.pack_value(vbf, inp["shapes"], inp["level_start"])
for _ in range(3):
    ops.ms_deform_attn_forward(vbf, inp["shapes"], inp["level_start"], loc, aw, 64)
    ops.ms_deform_attn_forward_packed(packed, vbf.shape, inp["shapes"], inp["level_start"], loc, aw)
torch.cuda.synchronize()
