#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_graph.py tests/test_gpu_dpj.py -m gpu -q -p no:cacheprovider --tb=short > gpurun_out/pytest_dist2c.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_dist2c.log
