#!/bin/bash
mkdir -p gpurun_out
REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 300 python tools/k_debug.py 2>&1 | tail -20
echo "== same with the generic (plain-load) GEMM kernel"
REST_B200_SPLIT_ORDER=0 REST_B200_FUSED_SPLITK=0 timeout -k 10 600 python tools/k_debug.py 1 2>&1 | tail -20
timeout -k 10 300 tools/micro/copy_bench > gpurun_out/copy_bench.txt 2>&1; tail -80 gpurun_out/copy_bench.txt
