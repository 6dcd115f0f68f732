#!/bin/bash
# Compile a scratch translation unit that instantiates a few kernels (seconds, no GPU) and print SASS statistics.
#   tools/sass_try.sh <scratch.cu> <kernel-name substring> [extra nvcc flags...]
SRC=$1; PAT=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I"$ROOT/mdqe_cvpr2023_b200/csrc" "$@" -Xptxas -v -cubin -o /tmp/sass_try.cubin "$SRC" 2>&1 | grep -E "error|warning|registers|spill" | grep -v "^$" | head -20
python "$ROOT/tools/sass_stats.py" /tmp/sass_try.cubin "$PAT" --top 32
