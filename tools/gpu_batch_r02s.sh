#!/bin/bash
mkdir -p gpurun_out
tools/micro/launch_bench | tee gpurun_out/launch_bench.txt
for i in 1 2 3; do timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -2; done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
