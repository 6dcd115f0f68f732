#!/bin/bash
# Same-box A/B of library builds on the encoder-shape kernels: tools/ab_libs.sh <tag> <lib1.so> <lib2.so> ...
# (paths relative to the repo root; each build runs tools/bwd_quick.py in its own process via MSDA_B200_LIB)
TAG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib (pass $rep)"
    MSDA_B200_LIB=$PWD/$lib python tools/bwd_quick.py ${BWDQ_ARGS} 2>&1 | grep -E "R50_360|swinl|R50_720 +local"
  done
done | tee gpurun_out/ab_${TAG}.txt
