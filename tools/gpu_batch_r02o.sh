#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_contractions.py tests/test_gpu_iajb.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_sk.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_sk.log
for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1; done | tee gpurun_out/k_timings_sk2.txt
timeout -k 10 300 python tools/sweep_e.py 2>&1 | grep '"n"' | cut -c1-500
for sk in 0 1; do
REST_B200_STREAMK=$sk timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/small_gemm_sk$sk.csv python tools/prof_small_gemm.py > /dev/null 2>&1
echo "== streamk=$sk kernel times (us)"; grep time_duration gpurun_out/small_gemm_sk$sk.csv | grep -v fill | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | cut -c1-60,100- | tail -24
done
