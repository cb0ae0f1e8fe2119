#!/bin/bash
# round 2, second pass (1 GPU): full GPU suite with the packed-M / block-list GEMM core, smoke, bench N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cut -c1-1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
