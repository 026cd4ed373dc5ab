#!/bin/bash
# ncu --set full of the DRR kernel for sweep cases given as args ("layout,view,sigma,order" ...)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_metrics.py -m gpu -q > gpurun_out/pytest_metrics.log 2>&1; tail -3 gpurun_out/pytest_metrics.log
for c in "$@"; do
  tag=$(echo $c | tr ',' '_')
  SWEEP_ONLY=$c timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_ -s 3 -c 1 -f \
      -o gpurun_out/prof_drr_$tag python scripts/sweep_layouts.py > gpurun_out/prof_drr_$tag.log 2>&1
  tail -2 gpurun_out/prof_drr_$tag.log
done
ls -la gpurun_out
