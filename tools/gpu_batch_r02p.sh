#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1; done | tee gpurun_out/k_timings_sk3.txt
timeout -k 10 300 python tools/sweep_e.py 2>&1 | grep '"n"' | cut -c1-420
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/small_gemm_sk1b.csv python tools/prof_small_gemm.py > /dev/null 2>&1
grep time_duration gpurun_out/small_gemm_sk1b.csv | grep -v fill | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | cut -c1-60,100- | tail -36
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/r02_hbm_kernels.json > gpurun_out/hbm_final.log 2>&1; echo "hbm rc=$?"
