#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_layout.py tests/test_gpu_contractions.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_layout.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_layout.log
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/hbm_256b.json > gpurun_out/hbm_256b.log 2>&1; echo "hbm rc=$?"; cat gpurun_out/hbm_256b.json
timeout -k 10 300 tools/micro/copy_bench > gpurun_out/copy_bench2.txt 2>&1; grep -E "mix32|cudaMemcpy" gpurun_out/copy_bench2.txt
timeout -k 10 300 python tools/prof_dp.py > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -k 10 300 ncu --metrics $M --clock-control none -c 12 --csv --log-file gpurun_out/dp_launches.csv python tools/prof_dp.py > gpurun_out/prof_dp.log 2>&1; echo "ncu dp rc=$?"
grep -E "gemv" gpurun_out/dp_launches.csv | grep time_duration | cut -d, -f5,12- | head -8
