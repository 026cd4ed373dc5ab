#!/bin/bash
# round 2, call AF (2 GPUs): bench.py after the fall-back change: the normal tile-sharded path and the forced fall-back to pose sharding
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29751 bench.py --gpus 2 --steps 10 --warmup 3 --no-weak > gpurun_out/r02_af_normal.log 2>&1; echo "normal exit $?"
XRC_BENCH_FORCE_IPC_FAIL=1 timeout 300 $TR --nproc-per-node 2 --master-port 29752 bench.py --gpus 2 --steps 10 --warmup 3 --no-weak > gpurun_out/r02_af_fallback.log 2>&1; echo "fallback exit $?"
python - <<PY
import json
for f in ('gpurun_out/r02_af_normal.log','gpurun_out/r02_af_fallback.log'):
    l=[x for x in open(f) if x.startswith('{')]
    if l:
        d=json.loads(l[-1]); print(f, 'step %.4f value %.1f e2e %.4f shard %s note %s' % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['shard'], d['config'].get('shard_note')))
    else:
        print(f, open(f).read()[-2500:])
PY
