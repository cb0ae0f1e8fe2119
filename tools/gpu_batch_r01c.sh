#!/bin/bash
# compute-sanitizer over every kernel family (incl. the ri3mo consumers) + ncu launch list of the timed step only
mkdir -p gpurun_out
timeout -k 10 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
timeout -k 10 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"
timeout -k 10 300 python -m pytest tests/test_gpu_iajb.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
