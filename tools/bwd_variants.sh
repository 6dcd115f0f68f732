#!/bin/bash
# A/B libraries of the backward kernel (gathers in flight per lane x resident CTAs promised to ptxas); run tools/bwd_quick.py with
# MSDA_B200_LIB pointing at each.  Build here (no GPU needed):  bash tools/bwd_variants.sh build ; on the GPU box: bash tools/bwd_variants.sh run
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared"
VARIANTS="${VARIANTS:-4:3 2:3 4:4 2:4 8:2 8:3}"
if [ "$1" = "build" ]; then
  for v in $VARIANTS; do
    b=${v%%:*}; m=${v##*:}
    (cd mdqe_cvpr2023_b200/csrc && nvcc $FLAGS -DMSDA_BWD_BATCH=$b -DMSDA_BWD_MINB=$m -o ../libmsda_b200_b${b}m${m}_dbg.so msda_api.cu mask_gemm.cu consumers.cu) &
  done
  wait
  ls -la mdqe_cvpr2023_b200/*_dbg.so
else
  echo "== default (batch 4, minb 1)"; python tools/bwd_quick.py 2>&1 | grep -E "R50_360|R50_720 +local +float32"
  for v in $VARIANTS; do
    b=${v%%:*}; m=${v##*:}
    echo "== batch $b, min CTAs/SM $m"
    MSDA_B200_LIB=$PWD/mdqe_cvpr2023_b200/libmsda_b200_b${b}m${m}_dbg.so python tools/bwd_quick.py 2>&1 | grep -E "R50_360|R50_720 +local +float32"
  done
fi
