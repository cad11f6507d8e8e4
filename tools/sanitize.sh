#!/bin/bash
# compute-sanitizer over a few small parity tests (memcheck: every kernel path of N = 512 / 1024 / 4096 incl. the
# TMA-staged ones; racecheck: shared-memory hazards of the fused accumulate + FFT kernels) -> gpurun_out/r2_sanitizer.txt
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
: > $OUT
for tool in memcheck synccheck initcheck racecheck; do
  echo "== compute-sanitizer --tool $tool  (pytest -k 'cfg1_burst or n512_pairs or fallback_paths or validation_and_state')" >> $OUT
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 python -m pytest tests/test_gpu_engine_parity.py -m gpu -q -x -k "cfg1_burst or n512_pairs or fallback_paths or validation_and_state" > /tmp/san_$tool.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Hazard|Invalid" /tmp/san_$tool.log | sort | uniq -c | head -20 >> $OUT
done
cat $OUT
