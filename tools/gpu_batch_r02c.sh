#!/bin/bash
# round 2, K-build profile: eigen-solver fix check, K timings (split order A/B), ncu launch list with pipe/DRAM metrics, FP64 probe capture
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_eig.py tests/test_gpu_contractions.py tests/test_gpu_layout.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_eig.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/pytest_eig.log
REST_B200_LAYOUT_TMA=0 timeout -k 10 300 python tools/hbm_probe.py gpurun_out/hbm_plain.json > gpurun_out/hbm_plain.log 2>&1; echo "hbm plain rc=$?"
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/hbm_tma.json > gpurun_out/hbm_tma.log 2>&1; echo "hbm tma rc=$?"; cat gpurun_out/hbm_plain.json gpurun_out/hbm_tma.json
for combo in "0 0" "1 0" "1 1"; do
  set -- $combo
  for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do
    REST_B200_SPLIT_ORDER=$1 REST_B200_FUSED_SPLITK=$2 timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1 | sed "s/^/split_order=$1 fused=$2 /"
  done
done | tee gpurun_out/k_timings.txt
M=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
timeout -k 10 600 ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/k_launches_C.csv python tools/prof_k.py 600 1700 60 > gpurun_out/prof_k_C.log 2>&1
echo "ncu K C rc=$?"
timeout -k 10 600 ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/k_launches_B.csv python tools/prof_k.py 264 720 21 > gpurun_out/prof_k_B.log 2>&1
echo "ncu K B rc=$?"
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:rb_gemm_tma_kernel -s 2 -c 2 -o gpurun_out/k_full_C -f python tools/prof_k.py 600 1700 60 > gpurun_out/prof_k_full.log 2>&1
echo "ncu full rc=$?"
timeout -k 10 300 ncu --set full --clock-control none -k regex:probe -c 2 -o gpurun_out/fp64_probe -f python -c "
import sys; sys.path.insert(0,'.')
from rest_tensors_b200.device import Context
c=Context(0); print(c.fp64_peak_probe(0), c.fp64_peak_probe(1))" > gpurun_out/prof_probe.log 2>&1
echo "ncu probe rc=$?"; tail -2 gpurun_out/prof_probe.log
ls -la gpurun_out/*.ncu-rep
