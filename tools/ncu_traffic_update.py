#!/usr/bin/env python
"""Refresh profiles/ncu_traffic.json (what bench.py quotes as `roofline.traffic`) from a tools/ncu_summary.py JSON of the current
kernels:   python tools/ncu_traffic_update.py profiles/r02zz_ncu_full.json "description of the capture"
Encoder-sized sampling kernels are the launches with the largest grid; wavefronts = pct x cycles x 148 SMs of the same capture."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, note = sys.argv[1], sys.argv[2]
    rows = json.load(open(src))
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    db = json.load(open(path))

    def biggest(prefix):
        cand = [r for r in rows if r["kernel"].replace("void ", "").startswith(prefix)]
        return max(cand, key=lambda r: r["grid"]) if cand else None

    def traffic(r):
        return int(round((r["dram_read_MB"] + r["dram_write_MB"]) * 1e6))

    def wavefronts(r):
        return int(round(r["l1_wavefront_pct"] / 100.0 * r["sm_cycles"] * 148))

    fwd, bwd = biggest("msda_fwd_fast2_kernel"), biggest("msda_bwd_fast2_kernel")
    if fwd:
        db["msda_fwd_fast2_kernel"] = traffic(fwd)
        db["msda_fwd_fast2_kernel_l1_wavefronts"] = wavefronts(fwd)
    if bwd:
        db["msda_bwd_fast2_kernel"] = traffic(bwd)
        db["msda_bwd_fast2_kernel_l1_wavefronts"] = wavefronts(bwd)
        db["msda_bwd_fast2_kernel_l2_red_sectors"] = int(bwd["l2_red_sectors"])
    for r in rows:
        k = r["kernel"].replace("void ", "")
        if k.startswith("mask_fwd_tc4_kernel<float, 0>"):
            db["mask_fwd_tc4_kernel"] = traffic(r)
        elif k.startswith("mask_fwd_tc4_kernel<float, 1>"):
            db["mask_fwd_tc4_kernel_transB"] = traffic(r)
        elif k.startswith("mask_grad_coeff_tc_kernel"):
            db["mask_grad_coeff_tc_kernel"] = traffic(r)
        elif k.startswith("gemm3x_kernel<0, 0>"):
            db["gemm3x_kernel_fwd_20400x256x256"] = traffic(r)
        elif k.startswith("gemm3x_kernel<0, 1>"):
            db["gemm3x_kernel_dgrad_20400x256x256"] = traffic(r)
        elif k.startswith("gemm3x_kernel<1, 1>"):
            db["gemm3x_kernel_wgrad_20400x256x256"] = traffic(r)
    if "mask_grad_coeff_tc_kernel" in db and "mask_fwd_tc4_kernel_transB" in db:
        db["mask_backward_tc"] = db["mask_grad_coeff_tc_kernel"] + db["mask_fwd_tc4_kernel_transB"]
    db["source"] = "%s (%s): ncu --set full --clock-control none, R50_ovis_360 shapes, fp32; cold L2 (tools/gpu_round2.sh)" % (os.path.relpath(src, ROOT), note)
    json.dump(db, open(path, "w"), indent=1)
    print(json.dumps({k: v for k, v in db.items() if isinstance(v, int)}, indent=1))


if __name__ == "__main__":
    main()
