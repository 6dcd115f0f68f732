#!/bin/bash
# Sweep NCCL settings for the multi-GPU step (whole-step CUDA graph with captured all-reduces).  Usage: bash tools/nccl_sweep.sh N TAG
N=${1:-2}; TAG=${2:-sweep}
OUT=gpurun_out/nccl_sweep_${TAG}_${N}gpu.txt; : > $OUT
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 20 --warmup 3 --no-e2e --no-other-configs 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('%-60s step %.4f ms  without all-reduce %.4f ms  value %.1f' % ('$*', d['ms_per_step'], d.get('ms_per_step_without_allreduce') or 0, d['value']))" | tee -a $OUT
}
run NCCL_NVLS_ENABLE=0
run NCCL_NVLS_ENABLE=0 NCCL_MAX_CTAS=16
run NCCL_NVLS_ENABLE=0 NCCL_MAX_CTAS=8
run NCCL_NVLS_ENABLE=0 NCCL_MAX_CTAS=4
run NCCL_NVLS_ENABLE=0 NCCL_MAX_CTAS=2
run NCCL_NVLS_ENABLE=1
run NCCL_NVLS_ENABLE=1 NCCL_MAX_CTAS=8
run NCCL_NVLS_ENABLE=1 NCCL_MAX_CTAS=4
run NCCL_NVLS_ENABLE=0 NCCL_MAX_CTAS=8 NCCL_PROTO=Simple
