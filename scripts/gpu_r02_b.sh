#!/bin/bash
# round 2, call B (2 GPUs): GPU suite (multi-device tests on two DISTINCT devices), sharded bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_b_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_2gpu.log
tail -15 gpurun_out/pytest_gpu_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist/sharded_device_check.py > gpurun_out/r02_sharded_device_check_2gpu.jsonl 2> gpurun_out/r02_sharded_device_check_2gpu.err; echo "check exit $?"
cat gpurun_out/r02_sharded_device_check_2gpu.jsonl | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_2gpu.log 2>&1; echo "bench2 exit $?"; tail -c 3500 gpurun_out/r02_bench_c2_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --workload c4 > gpurun_out/r02_bench_c4_2gpu.log 2>&1; echo "bench c4 exit $?"; tail -c 1200 gpurun_out/r02_bench_c4_2gpu.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02_bench_c2_b.log 2>&1; tail -c 1500 gpurun_out/r02_bench_c2_b.log
