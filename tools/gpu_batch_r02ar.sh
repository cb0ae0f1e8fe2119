#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 50 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok" gpurun_out/sanitize_$tool.log | head -3
  grep -E "hazard|Uninitialized|Invalid|Barrier error" gpurun_out/sanitize_$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -5
done
