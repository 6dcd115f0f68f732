import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import load_golden, module_from_golden, module_inputs, nerr
from mdqe_cvpr2023_b200 import MSDeformAttn
for name in ["module_spatial_pred", "module_spatial_grid", "module_temporal_grid"]:
    for tc in (True, False):
        z = load_golden(name)
        mod = module_from_golden(z, MSDeformAttn).cuda()
        mod.tc_linear = tc
        query, ref, inp, shapes, mask = module_inputs(z, "cuda")
        out = mod(query, ref, inp, shapes, mask)
        out.backward(torch.from_numpy(z["grad_out"]).cuda())
        print(name, "tc" if tc else "torch", "out %.2e gq %.2e gi %.2e" % (nerr(out, z["out"]), nerr(query.grad, z["grad_query"]), nerr(inp.grad, z["grad_input"])),
              {k: "%.1e" % nerr(p.grad, z["gp." + k]) for k, p in mod.named_parameters()}, "mask" , None if mask is None else int(mask.sum()))
