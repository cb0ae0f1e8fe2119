#!/bin/bash
mkdir -p gpurun_out
timeout 600 tools/micro/copy_bench > gpurun_out/copy_bench_fixed.txt 2>&1; echo "rc=$?"; grep -E "cudaMemcpy|stride32|mix32|stride16 U=8 hint=none" gpurun_out/copy_bench_fixed.txt
