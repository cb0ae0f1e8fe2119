#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:rb_gemm_tma_kernel -s 1 -c 1 \
    -o gpurun_out/thin_n8 -f python tools/prof_thin.py > gpurun_out/prof_thin.log 2>&1
echo "ncu thin rc=$?"; tail -3 gpurun_out/prof_thin.log
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:rb_gemm_tma_kernel -s 3 -c 1 \
    -o gpurun_out/full_n128 -f python tools/prof_thin.py > gpurun_out/prof_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/prof_full.log
ls -la gpurun_out/*.ncu-rep
