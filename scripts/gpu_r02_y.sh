#!/bin/bash
# round 2, call Y (N GPUs): exactly the driver's SCALE command at N = $1 (our arm, default flags)
N=${1:-4}
set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_driver_ours_${N}gpu.log 2>&1; echo "ours exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r02_driver_ours_${N}gpu.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("N=$N: step %.4f ms value %.1f | e2e %.4f ms %.1f | drr %.4f ms frac %.3f | weak %s | plan %s" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('weak'), d['config'].get('tile_plan')))
else:
    print(open('gpurun_out/r02_driver_ours_${N}gpu.log').read()[-3000:])
PY
