#!/bin/bash
# round 2, call L (1 GPU): two-phase patch_seqsum2_kernel -- bit-exactness tests, metric tests, 13-pose step + launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q > gpurun_out/pytest_metrics_l.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_metrics_l.log
timeout 600 python bench.py --batch 13 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_b13_l.log 2>&1; tail -c 1500 gpurun_out/r02_bench_c2_b13_l.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_b13_l.csv \
    python bench.py --batch 13 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_b13_l.log 2>&1
grep -E "seqsum|patch_kernel|grad_fast|finalize" gpurun_out/r02_launches_b13_l.csv | tail -8 | cut -c1-220
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_l.log 2>&1; tail -c 1500 gpurun_out/r02_bench_c2_l.log
