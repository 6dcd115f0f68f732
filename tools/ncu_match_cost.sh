#!/bin/bash
# ncu --set full of the tensor-core matcher-cost kernel (R50_ovis_360 shape) -> gpurun_out/prof_match_cost_tc_$1.ncu-rep + summary
TAG=${1:-r02}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_cost_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_match_cost_tc_${TAG} \
    python tools/consumers_bench.py > gpurun_out/ncu_match_cost_tc_${TAG}.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_match_cost_tc_${TAG}.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
want = ('gpu__time_duration.sum', 'sm__inst_executed.sum', 'sm__pipe_tensor_subpipe', 'sm__inst_executed_pipe_tensor', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct', 'smsp__inst_executed_pipe_xu', 'sm__pipe_fma_cycles_active.avg.pct', 'smsp__pcsamp_warps_issue_stalled')
for h, v in zip(hdr, vals):
    if any(h.startswith(w) for w in want): print(f'{h:90s} {v}')
" > gpurun_out/ncu_match_cost_tc_${TAG}.txt
head -60 gpurun_out/ncu_match_cost_tc_${TAG}.txt
