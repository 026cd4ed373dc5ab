#!/bin/bash
# round 2, call K (2 GPUs): tile-sharded objective (peer stores over NVLink): parity + bench against pose sharding
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_sharded_device.py -m gpu -x -q -rs > gpurun_out/pytest_sharded_2gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_sharded_2gpu.log
timeout 600 $TR --nproc-per-node 2 --master-port 29711 tests/dist/sharded_device_check.py > gpurun_out/r02_sharded_device_check_2gpu.jsonl 2> gpurun_out/r02_sharded_device_check_2gpu.err; echo "check exit $?"; grep -c true gpurun_out/r02_sharded_device_check_2gpu.jsonl; grep false gpurun_out/r02_sharded_device_check_2gpu.jsonl | head -5; tail -5 gpurun_out/r02_sharded_device_check_2gpu.err
for sh in tiles poses; do
  timeout 900 $TR --nproc-per-node 2 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --shard $sh > gpurun_out/r02_bench_c2_2gpu_$sh.log 2>&1; echo "bench $sh exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c2_2gpu_$sh.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("$sh: step %.4f ms value %.1f | e2e %.4f ms | drr %.4f ms frac %.3f | weak %s" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('weak',{}).get('ms_per_step')))
else:
    print(open('gpurun_out/r02_bench_c2_2gpu_$sh.log').read()[-3000:])
PY
done
