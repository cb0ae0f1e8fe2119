#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_dpj.py -m gpu -q -p no:cacheprovider --tb=short -x 2>&1 | tail -3
timeout -k 10 200 python tools/prof_dpj.py 2>&1 | tee gpurun_out/prof_dpj_220.txt | tail -7
