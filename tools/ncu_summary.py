#!/usr/bin/env python
"""Condense .ncu-rep files (ncu --set full) into one JSON + markdown table for profiles/.

    python tools/ncu_summary.py OUT_PREFIX rep1.ncu-rep [rep2.ncu-rep ...]
"""
import csv
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_sectors_srcunit_tex_op_red.sum": "l2_red_sectors",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefront_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum": "l1_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "l1_wavefronts_shared",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
}


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
    return x * scale[unit] if unit in scale else x


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"report": rep.split("/")[-1], "kernel": r[hdr.index("Kernel Name")][:110]}
        for i, h in enumerate(hdr):
            if h in KEYS:
                d[KEYS[h]] = to_float(r[i], units[i])
        if "dram_read_MB" in d and "dram_write_MB" in d:
            d["dram_traffic_MB"] = d["dram_read_MB"] + d["dram_write_MB"]
        res.append(d)
    return res


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    allk = []
    for rep in reps:
        allk += summarize(rep)
    json.dump(allk, open(prefix + ".json", "w"), indent=1)
    cols = ["duration_us", "dram_traffic_MB", "dram_pct", "l2_pct", "l1_wavefront_pct", "l1_hit_pct", "l2_hit_pct",
            "issue_active_pct", "warps_active_pct", "registers", "tensor_pipe_pct"]
    with open(prefix + ".md", "w") as f:
        f.write("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for d in allk:
            f.write("| " + d["kernel"].replace("|", "/") + " | " + " | ".join(
                (f"{d[c]:.1f}" if isinstance(d.get(c), float) else str(d.get(c, ""))) for c in cols) + " |\n")
    print(open(prefix + ".md").read())


if __name__ == "__main__":
    main()
