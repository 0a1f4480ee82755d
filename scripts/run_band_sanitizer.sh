#!/bin/bash
# compute-sanitizer passes over the banded kernels + contact kernel (run on the GPU box); log -> $1
LOG=${1:-gpurun_out/band_sanitizer.log}
: > $LOG
for tgt in "small 2" "world 2"; do
  for tool in memcheck synccheck racecheck; do
    echo "=== compute-sanitizer --tool $tool  scripts/sanitize_band_target.py $tgt" >> $LOG
    timeout 400 compute-sanitizer --tool $tool python scripts/sanitize_band_target.py $tgt 2>&1 | grep -E "COMPUTE-SANITIZER|done|SUMMARY|ERROR|hazard|Error" | head -20 >> $LOG
  done
done
