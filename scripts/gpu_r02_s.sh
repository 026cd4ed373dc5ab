#!/bin/bash
# round 2, call S (1 GPU): patch_kernel with the stride-1 specialisation / unrolled warp offsets / 4 CTAs per SM
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_fuzz.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_s.log
for slots in 3 4; do
XRC_PATCH_SLOTS=$slots timeout 600 python bench.py --batch 13 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_b13_s$slots.log 2>&1
XRC_PATCH_SLOTS=$slots timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_c2_s$slots.log 2>&1
python - <<PY
import json
for f in ['gpurun_out/r02_bench_c2_b13_s$slots.log','gpurun_out/r02_bench_c2_s$slots.log']:
    l=[x for x in open(f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, 'step', d['ms_per_step'], 'drr', d['roofline']['kernel_ms'], 'e2e', d['e2e']['ms_per_step'])
    else: print(open(f).read()[-1500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_b13_s.csv \
    python bench.py --batch 13 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_b13_s.log 2>&1
grep -E "seqsum|patch_kernel|grad_fast|finalize" gpurun_out/r02_launches_b13_s.csv | tail -4 | awk -F'","' '{print $5, $9, $NF}'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c2_s.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_c2_s.log 2>&1
grep -E "seqsum|patch_kernel|grad_fast|finalize" gpurun_out/r02_launches_c2_s.csv | tail -4 | awk -F'","' '{print $5, $9, $NF}'
