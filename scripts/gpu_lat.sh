#!/bin/bash
# Latency-regime iteration: GPU tests, latency script, ncu launch lists of the population-1 cases.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/latency.py > gpurun_out/latency.log 2>&1
cat gpurun_out/latency.log
for c in 192,grad-ncc,1,3 384,grad-ncc,1,3 768,grad-ncc,1,3 192,patch-grad-ncc,1,3; do
  tag=$(echo $c | tr ',' '_')
  LAT_ONLY=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/lat_$tag.csv \
      python scripts/latency.py > gpurun_out/lat_$tag.log 2>&1
done
