#!/bin/bash
# round 2, call P (8 GPUs): tile-sharded objective on 8 distinct devices: parity test + C2 strong scaling, tiles vs poses
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_sharded_device.py -m gpu -x -q -rs > gpurun_out/pytest_sharded_8gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_sharded_8gpu.log
for sh in tiles poses; do
  timeout 600 $TR --nproc-per-node 8 --master-port 29722 bench.py --gpus 8 --steps 20 --warmup 3 --shard $sh --no-cpu-baseline --no-weak > gpurun_out/r02_bench_c2_8gpu_$sh.log 2>&1; echo "bench $sh exit $?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/r02_bench_c2_8gpu_$sh.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print("$sh: step %.4f ms value %.1f | e2e %.4f ms | drr %.4f ms frac %.3f" % (d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))
else:
    print(open('gpurun_out/r02_bench_c2_8gpu_$sh.log').read()[-3000:])
PY
done
