#!/bin/bash
# One gpurun call: full GPU test suite, bench (config C, 1 GPU), thin-edge-tile A/B probe, config B bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
# gate: the GEMM sweeps with the thin-edge-tile loads; if they fail or hang, the rest of the batch runs with them off
timeout -k 10 400 python -m pytest tests/test_gpu_contractions.py -m gpu -q -x -p no:cacheprovider -k "dgemm or syrk or gemm" > gpurun_out/pytest_gate.log 2>&1
rc=$?; echo "gate rc=$rc"; tail -3 gpurun_out/pytest_gate.log
if [ $rc -ne 0 ]; then export REST_B200_THIN=0; echo "THIN LOADS DISABLED FOR THE REST OF THE BATCH"; fi
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json
timeout 300 python tools/thin_tile_probe.py > gpurun_out/thin.log 2>&1
echo "thin rc=$?"; tail -22 gpurun_out/thin.log
timeout 300 python bench.py --config B --no-cpu --no-e2e > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err
echo "benchB rc=$?"; cut -c1-400 gpurun_out/bench_B.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"
