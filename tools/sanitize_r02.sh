#!/bin/bash
# compute-sanitizer evidence for round 2 (run under gpurun on one B200)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; tail -6 gpurun_out/r02_sanitizer_$tool.log
done
