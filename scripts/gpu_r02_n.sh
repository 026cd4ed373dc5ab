#!/bin/bash
# round 2, call N (1 GPU): cluster form of patch_seqsum2_kernel -- bit-exactness, timing by cluster size, 13-pose step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q > gpurun_out/pytest_metrics_n.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_metrics_n.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_seqsum_probe2.csv python scripts/seqsum_probe2.py > gpurun_out/r02_seqsum_probe2.log 2>&1
tail -6 gpurun_out/r02_seqsum_probe2.log
grep -E "seqsum" gpurun_out/r02_seqsum_probe2.csv | awk -F'","' '{print $5, $9, $NF}' | tr -d '"'
timeout 600 python bench.py --batch 13 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_b13_n.log 2>&1; tail -c 300 gpurun_out/r02_bench_c2_b13_n.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:patch_seqsum2 -s 2 -c 1 -f -o gpurun_out/prof_seqsum2b \
    python bench.py --batch 13 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_seqsum2b.log 2>&1
