mkdir -p gpurun_out
SEL='merged_reductions or level_point or head_configs or grouped or decoder_shapes or chunk_sizes'
for tool in memcheck racecheck; do
  timeout 270 compute-sanitizer --tool $tool --error-exitcode 7 --log-file gpurun_out/sanitizer_${tool}_r01z.log \
      python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_${tool}_pytest_r01z.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/sanitizer_${tool}_pytest_r01z.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|Race" gpurun_out/sanitizer_${tool}_r01z.log | sort | uniq -c | head -8
done
