#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:rb_unpack_upper4_kernel|rb_pack_upper|rb_transpose4_kernel" -s 3 -c 3 -o gpurun_out/r02_layout_full -f python tools/prof_layout.py 8000 > gpurun_out/prof_layout.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_layout.log; ls -la gpurun_out/r02_layout_full.ncu-rep
