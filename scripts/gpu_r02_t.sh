#!/bin/bash
# round 2, call T (1 GPU): the whole GPU suite, smoke, the default bench line, and the round's ncu evidence
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_c2_final.log 2>&1
tail -c 600 gpurun_out/r02_bench_c2_final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_pax -s 10 -c 1 -f -o gpurun_out/r02_prof_drr \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_drr.log 2>&1
tail -2 gpurun_out/r02_prof_drr.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"patch_kernel|grad_fast_kernel|patch_seqsum" -s 6 -c 3 -f -o gpurun_out/r02_prof_sim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_prof_sim.log 2>&1
tail -2 gpurun_out/r02_prof_sim.log
