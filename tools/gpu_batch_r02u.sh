#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider --tb=short -k "stale or host_register or split_k or stream_k" > gpurun_out/pytest_r02u.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_r02u.log
timeout -k 10 600 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/initcheck.log 2>&1
echo "initcheck rc=$?"; grep -c "Uninitialized" gpurun_out/initcheck.log; grep -A12 "Uninitialized" gpurun_out/initcheck.log | grep -E "Uninitialized|at .*rb_|at .*at::|in .*\.cu" | sort | uniq -c | sort -rn | head -30; tail -3 gpurun_out/initcheck.log
