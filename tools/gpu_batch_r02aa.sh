#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/micro/size_sweep > gpurun_out/size_sweep.txt 2>&1; echo "rc=$?"; cat gpurun_out/size_sweep.txt
