#!/bin/bash
# compute-sanitizer over the small all-kernel workload: plain run, memcheck, racecheck (shared memory), synccheck
T=${1:-r2s}
mkdir -p gpurun_out
python tools/sanitize_small.py > gpurun_out/${T}_plain.log 2>&1; echo "plain exit=$?"; tail -2 gpurun_out/${T}_plain.log | cut -c1-400
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/${T}_$tool.log 2>&1; echo "$tool exit=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^\{" gpurun_out/${T}_$tool.log | head -8 | cut -c1-300
done
