#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_dpj.py -m gpu -q -p no:cacheprovider --tb=short -x > gpurun_out/pytest_dpj.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_dpj.log
timeout -k 10 200 python tools/prof_dpj.py 2>&1 | tee gpurun_out/prof_dpj.txt | tail -8
