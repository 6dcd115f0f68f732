#!/bin/bash
# compute-sanitizer memcheck + racecheck over the small-shape parity tests (SURVEY test matrix t6).
mkdir -p gpurun_out
SEL='golden_fixtures or head_configs or level_point or grouped or decoder_shapes or chunk_sizes'
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/sanitizer_${tool}.log \
      python -m pytest tests/test_msda_gpu.py tests/test_mask_gpu.py tests/test_consumers_gpu.py -m gpu -q -x \
      -k "$SEL or mask_golden or unbatched or (consumers and (golden or vs_oracle) and not 196 and not 300 and not 96)" > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Race" gpurun_out/sanitizer_${tool}.log | sort | uniq -c | head -8
done
