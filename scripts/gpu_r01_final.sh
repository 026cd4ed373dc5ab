#!/bin/bash
# One gpurun call: whole GPU suite, smoke, the bench (with the CPU leg), the reference arm, the ncu launch list of the
# bench command and full captures of the metric-stage kernels.  Logs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.log 2>&1
tail -1 gpurun_out/bench_c2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|patch_seqsum|grad_fast" -s 9 -c 3 -f -o gpurun_out/prof_sim3 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sim3.log 2>&1
ls -la gpurun_out/ | tail -8
