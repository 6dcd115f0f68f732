#!/usr/bin/env python
"""ncu target: one eager forward+backward of an MSDeformAttn module (default configuration) after two warm-up steps.
   usage: module_profile_target.py {encoder_self_attn,decoder_frame_attn,decoder_clip_attn}
   The launch list of the LAST step is what tools/launch_table.py --last-step prints."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mdqe_cvpr2023_b200.modules as M  # noqa: E402

PYR = [(48, 80), (24, 40), (12, 20), (6, 10)]
S = sum(h * w for h, w in PYR)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "encoder_self_attn"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    shapes = torch.tensor(PYR, device="cuda")
    if what == "encoder_self_attn":
        mod, q, x = M.MSDeformAttn(256, 4, 8, 4, pred_offsets=True, mode="spatial").cuda(), torch.randn(4, S, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")
    elif what == "decoder_frame_attn":
        mod, q, x = M.MSDeformAttn(256, 4, 8, 4, pred_offsets=False, mode="spatial").cuda(), torch.randn(4, 196, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")
    else:
        mod, q, x = M.MSDeformAttn(256, 4, 8, 4, n_frames=4, pred_offsets=False, mode="temporal").cuda(), torch.randn(1, 196, 256, device="cuda"), torch.randn(1, 4, S, 256, device="cuda")
    Q = q.shape[1]
    ref = torch.cat([torch.rand(q.shape[0], Q, 2, device="cuda"), torch.full((q.shape[0], Q, 2), 0.1, device="cuda")], -1)
    q.requires_grad_(True)
    x.requires_grad_(True)
    for i in range(3):
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("step%d" % i)
        torch.autograd.grad(mod(q, ref, x, shapes, None).sum(), (q, x) + tuple(mod.parameters()))
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()


if __name__ == "__main__":
    main()
