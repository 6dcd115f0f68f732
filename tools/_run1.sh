mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py tests/test_dropin_gpu.py -m gpu -q -x 2>&1 | tail -3
for opt in "" "--tiles"; do
timeout 200 python tools/linear_bench.py $opt 2>&1 | tail -6 | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['shape'][:30], {k[3:-9]: v for k, v in r.items() if k.startswith('tc_') and 'graph' in k})"
done
timeout 300 python tools/module_bench.py 2>&1 | grep -v "^reference CUDA" | tail -4 | cut -c1-120
