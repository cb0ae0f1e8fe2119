#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:rb_ri_dp_j_kernel" -s 2 -c 1 -o gpurun_out/r02_dpj_full -f python tools/prof_dpj_one.py > gpurun_out/prof_dpj_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_dpj_ncu.log; ls -la gpurun_out/r02_dpj_full.ncu-rep
