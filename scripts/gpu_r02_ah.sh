#!/bin/bash
# round 2, call AH (1 GPU): the state the round ends with -- whole GPU suite, smoke, the default bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c2_final3.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c2_final3.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('bench: step %.4f value %.1f e2e %.1f drr %.4f frac %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))
else: print(open('gpurun_out/r02_bench_c2_final3.log').read()[-1500:])
PY
