#!/bin/bash
mkdir -p gpurun_out
for tool in synccheck racecheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 50 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" gpurun_out/sanitize_$tool.log | head -3
  grep -A1 -E "Barrier error|hazard" gpurun_out/sanitize_$tool.log | grep " at " | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -5
done
timeout -k 10 300 python -m pytest tests/test_gpu_dpj.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -3
timeout -k 10 200 python tools/prof_dpj.py 2>&1 | tail -5
