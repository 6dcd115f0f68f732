#!/bin/bash
# Short GPU session for kernel iteration: msda parity tests, quick kernel sweep, ncu full on the main kernels.
TAG=${1:-q}
mkdir -p gpurun_out
echo "== pytest msda"; timeout 900 python -m pytest tests/test_msda_gpu.py -m gpu -q -x > gpurun_out/pytest_msda_${TAG}.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_msda_${TAG}.log
echo "== sweep"; timeout 600 python tools/kernel_bench.py --iters 15 --quick > gpurun_out/kb_${TAG}.log 2>&1; echo "rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/kernel_bench.json')):
    if 'dist' not in r: continue
    print(r['shape'], r['dist'], 'nerr_vs_refcuda', r.get('max_nerr_vs_refcuda'))
    for k,v in r.items():
        if isinstance(v,dict): print('   %-22s %8.1f us  %7.0f GB/s  %.3f'%(k, v['median_us'], v['GBps'], v['frac_hbm']))
PY
for what in enc_fwd enc_bwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 1 -c 1 -f -o gpurun_out/prof_${what}_${TAG} \
      python tools/profile_target.py $what > gpurun_out/ncu_${what}_${TAG}.log 2>&1; echo "ncu $what rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_fwd_tc2 -s 1 -c 1 -f -o gpurun_out/prof_mask_fwd_${TAG} \
    python tools/profile_target.py mask_fwd --dtype bf16 > gpurun_out/ncu_mask_fwd_${TAG}.log 2>&1; echo "ncu mask_fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_grad -s 2 -c 2 -f -o gpurun_out/prof_mask_bwd_${TAG} \
    python tools/profile_target.py mask_bwd > gpurun_out/ncu_mask_bwd_${TAG}.log 2>&1; echo "ncu mask_bwd rc=$?"
timeout 100 python tools/mask_debug.py > gpurun_out/mask_timeline_${TAG}.txt 2>&1
