#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_dpj.py -m gpu -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_dpj.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_dpj.log
timeout -k 10 300 python tools/prof_dpj_sweep.py 2>&1 | tee gpurun_out/prof_dpj_sweep3.txt | tail -8
timeout -k 10 200 python tools/prof_dpj.py 2>&1 | tee gpurun_out/prof_dpj.txt | tail -6
