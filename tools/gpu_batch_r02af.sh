#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/e2e_head_probe.py > gpurun_out/e2e_head_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/e2e_head_probe.txt | tail -14
