#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool initcheck --print-limit 5000 --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/initcheck.log 2>&1
echo "initcheck rc=$?"; grep -c "Uninitialized" gpurun_out/initcheck.log; grep -A3 "Uninitialized" gpurun_out/initcheck.log | grep -E " at " | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30; tail -3 gpurun_out/initcheck.log
