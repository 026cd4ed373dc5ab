#!/bin/bash
# round 2, call H (8 GPUs): strong scaling of the BASELINE population (bench.py under torchrun, N = 8 / 4), C4 sharded over
# (view, pose), the C5 pose-batch sweep at 1 / 2 / 4 / 8 GPUs, the sharded parity check on 8 distinct devices
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_h_gpus.txt; nproc >> gpurun_out/r02_h_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_8gpu.log 2>&1; echo "c2x8 exit $?"; grep '^{' gpurun_out/r02_bench_c2_8gpu.log | cut -c1-400
timeout 600 $TR --nproc-per-node 4 --master-port 29602 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_c2_4gpu.log 2>&1; echo "c2x4 exit $?"; grep '^{' gpurun_out/r02_bench_c2_4gpu.log | cut -c1-300
timeout 600 $TR --nproc-per-node 8 --master-port 29603 tests/dist/sharded_device_check.py > gpurun_out/r02_sharded_device_check_8gpu.jsonl 2> gpurun_out/r02_sharded_device_check_8gpu.err; echo "check exit $?"; tail -3 gpurun_out/r02_sharded_device_check_8gpu.jsonl
timeout 900 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --steps 5 --warmup 3 --workload c4 > gpurun_out/r02_bench_c4_8gpu.log 2>&1; echo "c4x8 exit $?"; grep '^{' gpurun_out/r02_bench_c4_8gpu.log | cut -c1-300
timeout 900 $TR --nproc-per-node 8 --master-port 29605 scripts/sweep_sharded.py > gpurun_out/r02_sweep_c5_8gpu.jsonl 2> gpurun_out/r02_sweep_c5_8gpu.err; echo "sweep8 exit $?"; tail -4 gpurun_out/r02_sweep_c5_8gpu.jsonl
timeout 900 $TR --nproc-per-node 4 --master-port 29606 scripts/sweep_sharded.py > gpurun_out/r02_sweep_c5_4gpu.jsonl 2> gpurun_out/r02_sweep_c5_4gpu.err; echo "sweep4 exit $?"; tail -2 gpurun_out/r02_sweep_c5_4gpu.jsonl
timeout 900 $TR --nproc-per-node 2 --master-port 29607 scripts/sweep_sharded.py > gpurun_out/r02_sweep_c5_2gpu.jsonl 2> gpurun_out/r02_sweep_c5_2gpu.err; echo "sweep2 exit $?"; tail -2 gpurun_out/r02_sweep_c5_2gpu.jsonl
timeout 900 python scripts/sweep_sharded.py > gpurun_out/r02_sweep_c5_1gpu.jsonl 2> gpurun_out/r02_sweep_c5_1gpu.err; echo "sweep1 exit $?"; tail -2 gpurun_out/r02_sweep_c5_1gpu.jsonl
