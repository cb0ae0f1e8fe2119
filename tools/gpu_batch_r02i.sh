#!/bin/bash
# full GPU suite on the restored GEMM + 256-bit layout kernels; HBM probe; bench
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/hbm_256.json > gpurun_out/hbm_256.log 2>&1; echo "hbm rc=$?"; cat gpurun_out/hbm_256.json
timeout -k 10 300 python tools/sweep_e.py > gpurun_out/sweep_e.log 2>&1; echo "sweep rc=$?"; tail -25 gpurun_out/sweep_e.log
timeout -k 10 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
