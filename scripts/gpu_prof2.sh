#!/bin/bash
# ncu evidence after the trimming change: launch lists (bench + pop-1 latency cases), full capture of the DRR kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
for c in 192,grad-ncc,1,3 768,grad-ncc,1,3 192,patch-grad-ncc,1,3; do
  tag=$(echo $c | tr ',' '_')
  LAT_ONLY=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/lat_$tag.csv \
      python scripts/latency.py > gpurun_out/lat_$tag.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:drr_pax -s 4 -c 1 -f -o gpurun_out/prof_drr_trim \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_drr_trim.log 2>&1
tail -3 gpurun_out/prof_drr_trim.log
ls -la gpurun_out/
