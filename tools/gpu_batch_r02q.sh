#!/bin/bash
# compute-sanitizer over every kernel family with the round-2 kernels; final K launch lists; final bench N=1
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitize_memcheck.log
timeout -k 10 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize_racecheck.log
timeout -k 10 900 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitize_target.py > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/sanitize_synccheck.log
MK=gpu__time_duration.sum,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_C.csv python tools/prof_k.py 600 1700 60 > gpurun_out/prof_k_C.log 2>&1; tail -1 gpurun_out/prof_k_C.log
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_B.csv python tools/prof_k.py 264 720 21 > gpurun_out/prof_k_B.log 2>&1; tail -1 gpurun_out/prof_k_B.log
timeout -k 10 600 ncu --metrics $MK --clock-control none -c 30 --csv --log-file gpurun_out/r02_k_launches_D.csv python tools/prof_k.py 1800 600 180 > gpurun_out/prof_k_D.log 2>&1; tail -1 gpurun_out/prof_k_D.log
for cfg in "600 1700 60" "264 720 21" "1800 600 180" "100 400 20"; do timeout -k 10 300 python tools/prof_k.py $cfg 2>&1 | tail -1; done | tee gpurun_out/r02_k_timings.txt
timeout -k 10 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench n1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout -k 10 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
