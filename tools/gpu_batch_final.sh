#!/bin/bash
# final 1-GPU verification: full GPU suite, smoke(), bench (config C), the reference arm, eigen probe, HBM-kernel ncu capture
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -10 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_n1.json
timeout -k 10 200 python tools/eig_probe.py > gpurun_out/eig_probe.log 2>&1
echo "eig probe rc=$?"; tail -3 gpurun_out/eig_probe.log
timeout -k 10 300 ncu --set full --clock-control none -k regex:rb_gemv -s 4 -c 2 -o gpurun_out/gemv_cfgC -f python tools/prof_dp.py > gpurun_out/prof_dp.log 2>&1
echo "ncu gemv rc=$?"; tail -2 gpurun_out/prof_dp.log
