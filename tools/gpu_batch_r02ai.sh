#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_small_probe.py 2>&1 | tee gpurun_out/e2e_small_probe.txt | tail -4
