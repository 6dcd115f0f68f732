#!/usr/bin/env python
"""Per-kernel timing sweep on one GPU: our kernels (variants / chunk sizes) next to the reference's own
CUDA extension (oracle/_ref, when it was built) at the BASELINE shapes.  CUDA events around each launch,
L2 flushed between launches by writing a 512 MB buffer.  Writes gpurun_out/kernel_bench.json.

    python tools/kernel_bench.py [--iters 20] [--quick]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mdqe_cvpr2023_b200 import _lib, ops  # noqa: E402
from tests.gpu_util import R50_360, R50_720, make_inputs, to_cuda  # noqa: E402


def timed(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return {"median_us": ts[len(ts) // 2], "min_us": ts[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    flush = torch.empty(128 * 1024 * 1024, device="cuda")
    ref = None
    try:
        from oracle import build_ref_cuda
        ref = build_ref_cuda.load()
    except Exception as e:  # noqa: BLE001
        print("reference CUDA extension unavailable:", e)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}

    shapes = [("enc_R50_360", 4, R50_360, 32, None), ("dec_R50_360", 4, R50_360, 32, 196),
              ("enc_R50_720", 4, R50_720, 32, None), ("enc_swinl_360", 3, R50_360, 24, None)]
    if args.quick:
        shapes = shapes[:2]
    results = []
    for name, N, pyr, D, Lq in shapes:
        for dist in ("local", "uniform"):
            cpu = make_inputs(N, pyr, 8, D, 4, Lq=Lq, dist=dist, seed=0)
            inp = to_cuda(cpu)
            S = inp["value"].shape[1]
            lq = inp["loc"].shape[1]
            smp = N * lq * 8 * 16
            fwd_b = 4 * (N * S * 8 * D + 3 * smp + N * lq * 8 * D)
            bwd_b = 4 * (2 * N * S * 8 * D + 6 * smp + N * lq * 8 * D)
            a = (inp["value"], inp["shapes"], inp["level_start"], inp["loc"], inp["aw"])
            row = {"shape": name, "dist": dist, "N": N, "S": S, "Lq": lq, "D": D, "fwd_bytes": fwd_b, "bwd_bytes": bwd_b}

            def rec(key, t, nbytes):
                t["GBps"] = nbytes / t["median_us"] / 1e3
                t["frac_hbm"] = t["GBps"] / peaks["hbm_gbs"]
                row[key] = t

            for chunk in ((0,) if args.quick else (0, 16, 32, 64, 128, 256)):
                _lib.set_option("chunk_pairs", chunk)
                rec(f"ours_fwd_chunk{chunk}", timed(lambda: ops.ms_deform_attn_forward(*a, 64), args.iters, flush), fwd_b)
                rec(f"ours_bwd_chunk{chunk}", timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), args.iters, flush), bwd_b)
            _lib.set_option("chunk_pairs", 0)
            _lib.set_option("fwd_variant", 3)
            rec("ours_fwd_lean", timed(lambda: ops.ms_deform_attn_forward(*a, 64), args.iters, flush), fwd_b)
            _lib.set_option("fwd_variant", 2)
            _lib.set_option("bwd_variant", 2)
            rec("ours_fwd_gen1", timed(lambda: ops.ms_deform_attn_forward(*a, 64), args.iters, flush), fwd_b)
            rec("ours_bwd_gen1", timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), args.iters, flush), bwd_b)
            _lib.set_option("fwd_variant", 1)
            _lib.set_option("bwd_variant", 1)
            rec("ours_generic_fwd", timed(lambda: ops.ms_deform_attn_forward(*a, 64), args.iters, flush), fwd_b)
            rec("ours_generic_bwd", timed(lambda: ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), args.iters, flush), bwd_b)
            _lib.set_option("fwd_variant", 0)
            _lib.set_option("bwd_variant", 0)
            # bf16 storage
            b16 = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in inp.items()}
            ab = (b16["value"], b16["shapes"], b16["level_start"], b16["loc"], b16["aw"])
            rec("ours_bf16_fwd", timed(lambda: ops.ms_deform_attn_forward(*ab, 64), args.iters, flush), fwd_b // 2)
            rec("ours_bf16_bwd", timed(lambda: ops.ms_deform_attn_backward(*ab, b16["grad_out"], 64), args.iters, flush), bwd_b // 2)
            if ref is not None:
                rec("refcuda_fwd", timed(lambda: ref.ms_deform_attn_forward(*a, 64), args.iters, flush), fwd_b)
                rec("refcuda_bwd", timed(lambda: ref.ms_deform_attn_backward(*a, inp["grad_out"], 64), args.iters, flush), bwd_b)
                # second parity witness
                o1, o2 = ops.ms_deform_attn_forward(*a, 64), ref.ms_deform_attn_forward(*a, 64)
                g1, g2 = ops.ms_deform_attn_backward(*a, inp["grad_out"], 64), ref.ms_deform_attn_backward(*a, inp["grad_out"], 64)
                row["max_nerr_vs_refcuda"] = max(float((x - y).abs().max() / y.abs().max()) for x, y in [(o1, o2)] + list(zip(g1, g2)))
            results.append(row)
            print(json.dumps(row), flush=True)

    # mask contraction: ours vs torch.einsum (cuBLAS) on the same GPU
    for Q, T, plane, K in ((196, 4, (96, 160), 32), (196, 4, (160, 288), 32), (300, 8, (96, 160), 32), (196, 3, (96, 160), 24)):
        coeff = torch.tanh(torch.randn(1, Q, K, device="cuda"))
        proto = torch.randn(1, K, T, *plane, device="cuda")
        n = T * plane[0] * plane[1]
        row = {"shape": f"mask_Q{Q}_K{K}_N{n}", "bytes_f32": 4 * (Q * K + K * n + Q * n), "flops": 2 * Q * K * n}
        for variant in (0, 1):
            _lib.set_option("mask_variant", variant)
            t = timed(lambda: ops.mask_logits_forward(coeff, proto), args.iters, flush)
            t["GBps"] = row["bytes_f32"] / t["median_us"] / 1e3
            t["tflops"] = row["flops"] / t["median_us"] / 1e6
            row[f"ours_f32_variant{variant}"] = t
            c16, p16 = coeff.bfloat16(), proto.bfloat16()
            t = timed(lambda: ops.mask_logits_forward(c16, p16), args.iters, flush)
            t["GBps"] = row["bytes_f32"] / 2 / t["median_us"] / 1e3
            t["tflops"] = row["flops"] / t["median_us"] / 1e6
            row[f"ours_bf16_variant{variant}"] = t
        _lib.set_option("mask_variant", 0)
        t = timed(lambda: torch.einsum("bqm,bmthw->bqthw", coeff, proto), args.iters, flush)
        t["GBps"] = row["bytes_f32"] / t["median_us"] / 1e3
        row["torch_einsum_f32"] = t
        t = timed(lambda: torch.einsum("bqm,bmthw->bqthw", c16, p16), args.iters, flush)
        t["GBps"] = row["bytes_f32"] / 2 / t["median_us"] / 1e3
        row["torch_einsum_bf16"] = t
        go = torch.randn(1, Q, T, *plane, device="cuda")
        t = timed(lambda: ops.mask_logits_backward(coeff, proto, go), args.iters, flush)
        row["ours_bwd_f32"] = t
        results.append(row)
        print(json.dumps(row), flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
