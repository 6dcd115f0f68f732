#!/usr/bin/env python
"""Where the HOST time of an eager module step goes (cProfile over eager forward+backward of the three module kinds).
The module-level step is host-bound when it is not replayed as a CUDA graph (bench.py module_arm: eager vs graph)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mdqe_cvpr2023_b200.modules as M  # noqa: E402

PYR = [(48, 80), (24, 40), (12, 20), (6, 10)]
S = sum(h * w for h, w in PYR)


def main():
    torch.manual_seed(0)
    shapes = torch.tensor(PYR, device="cuda")
    mods = [(M.MSDeformAttn(256, 4, 8, 4, pred_offsets=True, mode="spatial").cuda(), torch.randn(4, S, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")),
            (M.MSDeformAttn(256, 4, 8, 4, pred_offsets=False, mode="spatial").cuda(), torch.randn(4, 196, 256, device="cuda"), torch.randn(4, S, 256, device="cuda")),
            (M.MSDeformAttn(256, 4, 8, 4, n_frames=4, pred_offsets=False, mode="temporal").cuda(), torch.randn(1, 196, 256, device="cuda"), torch.randn(1, 4, S, 256, device="cuda"))]
    refs = [torch.cat([torch.rand(q.shape[0], q.shape[1], 2, device="cuda"), torch.full((q.shape[0], q.shape[1], 2), 0.1, device="cuda")], -1) for _, q, _ in mods]
    for _, q, x in mods:
        q.requires_grad_(True)
        x.requires_grad_(True)

    def step():
        for (mod, q, x), ref in zip(mods, refs):
            torch.autograd.grad(mod(q, ref, x, shapes, None).sum(), (q, x) + tuple(mod.parameters()))

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("3 modules fwd+bwd eager: host issue %.0f us, incl. GPU drain %.0f us per step" % (t_issue / n * 1e6, t_all / n * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        step()
    pr.disable()
    torch.cuda.synchronize()
    out = io.StringIO()
    st = pstats.Stats(pr, stream=out)
    st.sort_stats("tottime").print_stats(28)
    print(out.getvalue()[:6000])
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(30)
    print(out.getvalue()[:6000])


if __name__ == "__main__":
    main()
