#!/bin/bash
# Multi-GPU step time with the gradient all-reduce done by NCCL or by msda_allreduce_f32 at different CTA counts.
# Usage: bash tools/allreduce_step_sweep.sh N TAG
N=${1:-2}; TAG=${2:-sweep}
OUT=gpurun_out/allreduce_step_${TAG}_${N}gpu.txt; : > $OUT
run() {
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-other-configs "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-44s step %.4f ms  without all-reduce %.4f ms  value %.1f  %s' % ('$*', d['ms_per_step'], d.get('ms_per_step_without_allreduce') or 0, d['value'], (d.get('allreduce') or '')[:60]))" | tee -a $OUT
}
run --allreduce-impl nccl
for c in ${CTAS:-4 8}; do run --allreduce-impl peer --allreduce-ctas $c; done
