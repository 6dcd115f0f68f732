#!/bin/bash
# Quick ncu counters (no --set full) of one profile_target call under different library options.
# Usage: bash tools/ncu_quick.sh TAG enc_bwd "pair_map=1" "pair_map=2" ...
TAG=$1; WHAT=$2; shift 2
M=gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
mkdir -p gpurun_out
for OPT in "$@"; do
  F=gpurun_out/ncuq_${TAG}_${WHAT}_$(echo $OPT | tr ',=' '__').csv
  timeout 300 ncu --metrics $M --clock-control none -k regex:msda_ -s 1 -c 1 --csv --log-file $F python tools/profile_target.py $WHAT --opt "$OPT" > /dev/null 2>&1
  echo "== $WHAT $OPT"; python - "$F" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ni, vi, ki = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name")
for r in rows[1:]:
    print(f"  {r[ni]:75s} {r[vi]}")
PY
done
