#!/bin/bash
# launch list of the timed bench step with the final build of round 2 (kernel names of the final build)
mkdir -p gpurun_out
M=gpu__time_duration.sum
timeout -k 10 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras --no-parity > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/r02_launches_final.csv
