#!/bin/bash
# One gpurun call: full GPU test suite, bench (config C, 1 GPU), config B bench, ncu launch list of the bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
timeout -k 10 300 python bench.py --config B --no-cpu --no-e2e > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err
echo "benchB rc=$?"; cut -c1-300 gpurun_out/bench_B.json
if [ "$1" = "ncu" ]; then
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"
fi
