#!/bin/bash
# One GPU session: parity tests, smoke, bench, kernel sweep, ncu launch list + full capture.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_${TAG}.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_${TAG}.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_${TAG}.log
echo "== bench local"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "== bench uniform"; timeout 600 python bench.py --steps 20 --warmup 3 --dist uniform --no-e2e --no-cpu-baseline > gpurun_out/bench_uniform_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; tail -c 1500 gpurun_out/bench_uniform_${TAG}.json
echo "== kernel sweep"; timeout 900 python tools/kernel_bench.py --iters 15 > gpurun_out/kernel_bench_${TAG}.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/kernel_bench_${TAG}.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_launch_${TAG}.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (bwd, fwd)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_fast -s 9 -c 3 -f -o gpurun_out/prof_bwd_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --layers 1 > gpurun_out/ncu_bwd_${TAG}.log 2>&1; echo "ncu bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_fast -s 9 -c 3 -f -o gpurun_out/prof_fwd_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --layers 1 > gpurun_out/ncu_fwd_${TAG}.log 2>&1; echo "ncu fwd rc=$?"
echo "== mask head (fp32 tensor-core kernels): breakdown, timeline, ncu"
timeout 200 python tools/mask_bwd_breakdown.py > gpurun_out/mask_breakdown_${TAG}.txt 2>&1; cat gpurun_out/mask_breakdown_${TAG}.txt | tail -4
timeout 100 python tools/mask_bwd_timeline.py > gpurun_out/mask_timeline_${TAG}.txt 2>&1
timeout 100 python tools/mask_keepraw_check.py > gpurun_out/mask_keepraw_${TAG}.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mask_ -s 3 -c 3 -f -o gpurun_out/prof_mask_${TAG} \
    python tools/mask_profile_target.py > gpurun_out/ncu_mask_${TAG}.log 2>&1; echo "ncu mask rc=$?"
echo "== callers either side of the path (csrc/consumers.cu): timing vs the reference's torch ops, ncu of the matcher-cost kernel"
timeout 300 python tools/consumers_bench.py > gpurun_out/consumers_bench_${TAG}.log 2>&1; tail -8 gpurun_out/consumers_bench_${TAG}.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_cost_kernel -s 2 -c 1 -f -o gpurun_out/prof_match_cost_${TAG} \
    python tools/consumers_bench.py > gpurun_out/ncu_match_cost_${TAG}.log 2>&1; echo "ncu match_cost rc=$?"
echo "== encoder-shape fwd / bwd quick table"; timeout 300 python tools/bwd_quick.py > gpurun_out/bwd_quick_${TAG}.log 2>&1; tail -12 gpurun_out/bwd_quick_${TAG}.log
ls -la gpurun_out | tail -20
echo "== backward merge on/off, decoder-sized calls (cold)"; timeout 200 python tools/merge_ab.py > gpurun_out/merge_ab_${TAG}.log 2>&1; grep float32 gpurun_out/merge_ab_${TAG}.log
timeout 200 python tools/small_calls_ab.py > gpurun_out/small_calls_${TAG}.log 2>&1; tail -2 gpurun_out/small_calls_${TAG}.log
timeout 200 python tools/module_bench.py > gpurun_out/module_bench_${TAG}.log 2>&1; tail -3 gpurun_out/module_bench_${TAG}.log
