#!/bin/bash
# round 2 profiles: ERIFold4 + layout tests, launch list of the timed bench step, ncu --set full of the ao2mo GEMM and the d_P GEMV,
# K-build launch lists (configs C and B), final HBM probe
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_eri.py tests/test_gpu_layout.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
M=gpu__time_duration.sum
timeout -k 10 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-extras --no-parity > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list rc=$?"
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:rb_gemm_tma_kernel -s 2 -c 2 -o gpurun_out/r02_ao2mo_full -f python tools/prof_ao2mo.py > gpurun_out/prof_ao2mo.log 2>&1
echo "ncu ao2mo full rc=$?"
timeout -k 10 600 ncu --set full --clock-control none -k regex:rb_gemv -s 4 -c 4 -o gpurun_out/r02_dp_full -f python tools/prof_dp.py > gpurun_out/prof_dp_full.log 2>&1
echo "ncu dp full rc=$?"
MK=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_C.csv python tools/prof_k.py 600 1700 60 > gpurun_out/prof_k_C.log 2>&1; tail -1 gpurun_out/prof_k_C.log
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_B.csv python tools/prof_k.py 264 720 21 > gpurun_out/prof_k_B.log 2>&1; tail -1 gpurun_out/prof_k_B.log
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_D.csv python tools/prof_k.py 1800 600 180 > gpurun_out/prof_k_D.log 2>&1; tail -1 gpurun_out/prof_k_D.log
for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1; done | tee gpurun_out/r02_k_timings.txt
timeout -k 10 300 python tools/hbm_probe.py gpurun_out/r02_hbm_kernels.json > gpurun_out/hbm_final.log 2>&1; echo "hbm rc=$?"
MH=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -k 10 600 ncu --metrics $MH --clock-control none -k regex:"unpack|pack_upper|transpose4|copy_flat4|copy3d|axpy4" -c 60 --csv --log-file gpurun_out/r02_hbm_launches.csv python tools/hbm_probe.py > /dev/null 2>&1; echo "ncu hbm rc=$?"
ls -la gpurun_out/*.ncu-rep
