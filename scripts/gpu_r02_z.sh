#!/bin/bash
# round 2, call Z (8 GPUs): the driver's SCALE command at N = 8 (default flags, with the weak figure), then C4 tile-sharded
set -x
mkdir -p gpurun_out
bash scripts/gpu_r02_y.sh 8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-weak > gpurun_out/r02_bench_c4_8gpu_tiles.log 2>&1; echo "c4 exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c4_8gpu_tiles.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("c4 N=8: step %.4f ms value %.1f | e2e %.4f ms %.1f | drr %.4f ms frac %.3f | plan %s" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config'].get('tile_plan')))
else:
    print(open('gpurun_out/r02_bench_c4_8gpu_tiles.log').read()[-3000:])
PY
