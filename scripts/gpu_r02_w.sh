#!/bin/bash
# round 2, call W (1 GPU): L2 prefetch hint in the DRR loop for the few-pose regime (XRC_PAX_PF = samples ahead)
set -x
mkdir -p gpurun_out
for pf in 0 16 32 64; do
  XRC_PAX_PF=$pf timeout 300 python bench.py --workload c3-768-pop1 --steps 50 --no-cpu-baseline > gpurun_out/r02_pf${pf}_768.log 2>&1
  XRC_PAX_PF=$pf timeout 300 python bench.py --workload c5 --batch 1 --steps 10 --no-cpu-baseline > gpurun_out/r02_pf${pf}_c5b1.log 2>&1
  XRC_PAX_PF=$pf XRC_PAX_PF_PROJS=100 timeout 300 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_pf${pf}_c2.log 2>&1
  python - <<PY
import json
for f in ['gpurun_out/r02_pf${pf}_768.log','gpurun_out/r02_pf${pf}_c5b1.log','gpurun_out/r02_pf${pf}_c2.log']:
    l=[x for x in open(f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print('pf $pf', f.split('_')[-1], 'step', round(d['ms_per_step'],4), 'drr', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4))
    else: print(open(f).read()[-800:])
PY
done
XRC_PAX_PF=32 timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_drr.py -m gpu -x -q 2>&1 | tail -3
