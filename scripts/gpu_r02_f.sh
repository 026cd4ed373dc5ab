#!/bin/bash
# round 2, call F: small-batch (13 poses = the 8-GPU share) DRR kernel variants; C5 / C2 resident volume after the stack fixes
set -x
mkdir -p gpurun_out
for o in 0 2 18 26; do
  timeout 300 python bench.py --batch 13 --steps 20 --no-cpu-baseline --order $o > gpurun_out/r02_b13_order$o.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02_b13_order$o.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("order $o: step %.4f ms  drr %.4f ms  e2e %.4f ms" % (d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']))
else:
    print("order $o failed"); print(open('gpurun_out/r02_b13_order$o.log').read()[-1500:])
PY
done
timeout 900 python bench.py --workload c5 --batch 16 --steps 4 --no-cpu-baseline > gpurun_out/r02_bench_c5_b.log 2>&1; tail -c 400 gpurun_out/r02_bench_c5_b.log
grep -o '"volume_bytes_resident": [0-9]*' gpurun_out/r02_bench_c5_b.log
