#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_eig.py -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/pytest_eig.log 2>&1
echo "pytest eig rc=$?"; tail -40 gpurun_out/pytest_eig.log
timeout -k 10 300 python tools/eig_probe.py > gpurun_out/eig_probe.log 2>&1
echo "eig probe rc=$?"; tail -5 gpurun_out/eig_probe.log
