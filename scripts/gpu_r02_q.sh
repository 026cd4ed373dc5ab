#!/bin/bash
# round 2, call Q (N GPUs, N = $1): NCCL-free tile step (xchg kernel) + clock-balanced tile plan: parity + bench by mode
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_sharded_device.py -m gpu -x -q -rs > gpurun_out/pytest_sharded_${N}gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_sharded_${N}gpu.log
timeout 600 $TR --nproc-per-node $N --master-port 29711 tests/dist/sharded_device_check.py > gpurun_out/r02_sharded_device_check_${N}gpu.jsonl 2> gpurun_out/r02_sharded_device_check_${N}gpu.err; echo "check exit $?"; grep -c true gpurun_out/r02_sharded_device_check_${N}gpu.jsonl; grep false gpurun_out/r02_sharded_device_check_${N}gpu.jsonl | head -5; tail -5 gpurun_out/r02_sharded_device_check_${N}gpu.err
for sh in tiles tiles-nccl poses; do
  timeout 600 $TR --nproc-per-node $N --master-port 29712 bench.py --gpus $N --steps 20 --warmup 3 --shard $sh --no-cpu-baseline --no-weak > gpurun_out/r02_bench_c2_${N}gpu_$sh.log 2>&1; echo "bench $sh exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c2_${N}gpu_$sh.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("$sh: step %.4f ms value %.1f | e2e %.4f ms | drr %.4f ms frac %.3f | plan %s" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config'].get('tile_plan')))
else:
    print(open('gpurun_out/r02_bench_c2_${N}gpu_$sh.log').read()[-3000:])
PY
done
