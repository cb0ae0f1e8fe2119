#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_dpj.py tests/test_gpu_graph.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -4
for tool in synccheck racecheck; do
  timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | head -2
done
