#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/prof_dpj_sweep.py 2>&1 | tee gpurun_out/prof_dpj_sweep.txt | tail -14
