#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small mixed batch (every tier, every window kind).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
