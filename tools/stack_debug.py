#!/usr/bin/env python
"""Diagnostic for tests/test_stack_gpu.py: gradient error of the CUDA stack against the CPU/oracle stack for the shipped
configuration and with the Linear layers / prologue switched back to torch (are the outliers pixel-centre flips?)."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mdqe_cvpr2023_b200.modules as M  # noqa: E402
from tests.helpers import OracleMSDAFunction, nerr  # noqa: E402
from tests.test_stack_gpu import DIM, PYRAMID, Stack, _inputs  # noqa: E402

B, T, Q = 2, 3, 50
torch.manual_seed(0)
stack_cpu = Stack(M.MSDeformAttn, T)
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for p in stack_cpu.parameters():
        p.add_(0.02 * torch.randn(p.shape, generator=g))
inp = _inputs(B, T, Q, g)
names = ("src", "q_box", "q_inst")
w = [torch.randn(s, generator=g) for s in ((B * T, sum(h * w for h, w in PYRAMID), DIM), (B * T, Q, DIM), (B, Q, DIM))]


def run_gpu(tc, fused):
    st = copy.deepcopy(stack_cpu).cuda()
    for m in st.modules():
        if isinstance(m, M.MSDeformAttn):
            m.fused_prologue, m.tc_linear = fused, tc
    dev = {k: (v.cuda() if v is not None else None) for k, v in inp.items()}
    for k in names:
        dev[k].requires_grad_(True)
    outs = st(**dev)
    sum((o * wi.cuda()).sum() for o, wi in zip(outs, w)).backward()
    return outs, {k: dev[k].grad for k in names}, dict((k, p.grad) for k, p in st.named_parameters())


res = {(tc, fu): run_gpu(tc, fu) for tc in (True, False) for fu in (True, False)}
orig = M.MSDeformAttnFunction
M.MSDeformAttnFunction = OracleMSDAFunction
M.ops.grouped_supported = lambda *a: False
for m in stack_cpu.modules():
    if isinstance(m, M.MSDeformAttn):
        m.fused_prologue, m.tc_linear = False, False
cpu = {k: (v.clone() if v is not None else None) for k, v in inp.items()}
for k in names:
    cpu[k].requires_grad_(True)
outs_cpu = stack_cpu(**cpu)
sum((o * wi).sum() for o, wi in zip(outs_cpu, w)).backward()
pg = dict((k, p.grad) for k, p in stack_cpu.named_parameters())
for key, (outs, gin, gp) in res.items():
    print("tc_linear=%s fused=%s" % key)
    print("   out nerr", [f"{nerr(a, b):.1e}" for a, b in zip(outs, outs_cpu)])
    for k in names:
        a, b = gin[k].double().cpu(), cpu[k].grad.double()
        d = (a - b).abs()
        rows = (d.amax(-1) > 1e-3 * b.abs().max()).sum().item()
        print(f"   grad {k}: max-norm {nerr(a, b):.1e}  rel-L2 {float((a - b).norm() / b.norm()):.1e}  rows off {rows} of {d.shape[0] * d.shape[1]}")
    worst = sorted(((nerr(gp[k], pg[k]), k) for k in pg), reverse=True)[:4]
    print("   worst param grads", [(f"{v:.1e}", k) for v, k in worst])
