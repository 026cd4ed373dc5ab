#!/bin/bash
# round 2, call A (1 GPU): GPU suite, smoke, one bench line per BASELINE config through the new bench.py
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 300 python bench.py --workload small --steps 3 --no-cpu-baseline > gpurun_out/r02_bench_small.log 2>&1; echo "small exit $?"
tail -c 1500 gpurun_out/r02_bench_small.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/r02_bench_c2.log 2>&1; tail -c 3000 gpurun_out/r02_bench_c2.log
for wl in c2-oblique c2-coarse c3-192 c3-768-pop1 c1; do
  timeout 600 python bench.py --workload $wl --steps 10 > gpurun_out/r02_bench_$wl.log 2>&1; echo "$wl exit $?"; tail -c 600 gpurun_out/r02_bench_$wl.log
done
timeout 900 python bench.py --workload c4 --steps 5 > gpurun_out/r02_bench_c4.log 2>&1; echo "c4 exit $?"; tail -c 600 gpurun_out/r02_bench_c4.log
timeout 900 python bench.py --workload c5 --batch 16 --steps 4 > gpurun_out/r02_bench_c5.log 2>&1; echo "c5 exit $?"; tail -c 600 gpurun_out/r02_bench_c5.log
ls -la gpurun_out | tail -15
