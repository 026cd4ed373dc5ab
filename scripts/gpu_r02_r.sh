#!/bin/bash
# round 2, call R (8 GPUs): C2 strong scaling with the NCCL-free tile step, with and without clock feedback on the tile plan
N=8
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29711 tests/dist/sharded_device_check.py > gpurun_out/r02_sharded_device_check_${N}gpu.jsonl 2> gpurun_out/r02_sharded_device_check_${N}gpu.err; echo "check exit $?"; grep -c true gpurun_out/r02_sharded_device_check_${N}gpu.jsonl; grep false gpurun_out/r02_sharded_device_check_${N}gpu.jsonl | head -5; tail -5 gpurun_out/r02_sharded_device_check_${N}gpu.err
for bal in 3 0; do
  timeout 600 $TR --nproc-per-node $N --master-port 29712 bench.py --gpus $N --steps 20 --warmup 3 --shard tiles --balance $bal --no-cpu-baseline --no-weak > gpurun_out/r02_bench_c2_${N}gpu_tiles_bal$bal.log 2>&1; echo "bench bal $bal exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c2_${N}gpu_tiles_bal$bal.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("bal $bal: step %.4f ms value %.1f | e2e %.4f ms | drr %.4f ms frac %.3f | plan %s" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config'].get('tile_plan')))
else:
    print(open('gpurun_out/r02_bench_c2_${N}gpu_tiles_bal$bal.log').read()[-3000:])
PY
done
