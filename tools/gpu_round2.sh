#!/bin/bash
# Round-2 GPU session: parity tests, smoke, bench, ncu launch list + full captures of the sampling kernels, sanitizer.
# Usage (repo root on the GPU box): bash tools/gpu_round2.sh TAG
TAG=${1:-r02z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_${TAG}.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_${TAG}.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 400 gpurun_out/bench_ref_${TAG}.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-other-configs > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (bwd, fwd)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_fast -s 9 -c 3 -f -o gpurun_out/prof_bwd_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-other-configs --layers 1 > gpurun_out/ncu_bwd_${TAG}.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_fast -s 9 -c 3 -f -o gpurun_out/prof_fwd_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-other-configs --layers 1 > gpurun_out/ncu_fwd_${TAG}.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:mask_ -s 3 -c 3 -f -o gpurun_out/prof_mask_${TAG} \
    python tools/mask_profile_target.py > gpurun_out/ncu_mask_${TAG}.log 2>&1; echo "ncu mask rc=$?"
for what in lin_fwd lin_dgrad lin_wgrad; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm3x -s 1 -c 1 -f -o gpurun_out/prof_${what}_${TAG} \
      python tools/profile_target.py $what > gpurun_out/ncu_${what}_${TAG}.log 2>&1; echo "ncu $what rc=$?"
done
python tools/ncu_summary.py gpurun_out/ncu_full_${TAG} gpurun_out/prof_bwd_${TAG}.ncu-rep gpurun_out/prof_fwd_${TAG}.ncu-rep gpurun_out/prof_mask_${TAG}.ncu-rep \
    gpurun_out/prof_lin_fwd_${TAG}.ncu-rep gpurun_out/prof_lin_dgrad_${TAG}.ncu-rep gpurun_out/prof_lin_wgrad_${TAG}.ncu-rep > /dev/null 2>&1; echo "summary rc=$?"
echo "== module launch lists"
for w in encoder_self_attn decoder_frame_attn decoder_clip_attn; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/module_launches_${w}_${TAG}.csv python tools/module_profile_target.py $w > /dev/null 2>&1
  python tools/launch_table.py gpurun_out/module_launches_${w}_${TAG}.csv --every 3 > gpurun_out/module_launches_${w}_${TAG}.txt; tail -1 gpurun_out/module_launches_${w}_${TAG}.txt
done
echo "== sanitizer (new kernels + sampling kernels)"
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/sanitizer_${tool}_${TAG}.log \
      python -m pytest tests/test_msda_gpu.py tests/test_consumers_gpu.py -m gpu -q -x \
      -k "merged_reductions or level_point or head_configs or grouped or decoder_shapes or chunk_sizes or (match_cost and (golden or vs_oracle) and not 196 and not 300)" > gpurun_out/sanitizer_${tool}_pytest_${TAG}.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tool}_pytest_${TAG}.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_${tool}_${TAG}.log | sort | uniq -c | head -4
done
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/sanitizer_memcheck_linear_${TAG}.log \
    python -m pytest tests/test_linear_gpu.py -m gpu -q -x -k "not 20400 and not 15300 and not flags_clean" > gpurun_out/sanitizer_memcheck_linear_pytest_${TAG}.log 2>&1
echo "memcheck linear rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_linear_pytest_${TAG}.log; grep -E "ERROR SUMMARY" gpurun_out/sanitizer_memcheck_linear_${TAG}.log | sort | uniq -c | head -3
echo "== kernel sweep / consumers / module"; timeout 600 python tools/kernel_bench.py --iters 15 > gpurun_out/kernel_bench_${TAG}.log 2>&1; tail -2 gpurun_out/kernel_bench_${TAG}.log | cut -c1-300
timeout 200 python tools/consumers_bench.py > gpurun_out/consumers_bench_${TAG}.log 2>&1; tail -4 gpurun_out/consumers_bench_${TAG}.log | cut -c1-200
timeout 300 python tools/module_bench.py > gpurun_out/module_bench_${TAG}.log 2>&1; tail -3 gpurun_out/module_bench_${TAG}.log | cut -c1-160; cp gpurun_out/module_bench.json gpurun_out/module_bench_${TAG}.json
timeout 200 python tools/linear_bench.py > gpurun_out/linear_bench_${TAG}.log 2>&1; cp gpurun_out/linear_bench.json gpurun_out/linear_bench_${TAG}.json
ls gpurun_out | grep ${TAG}
