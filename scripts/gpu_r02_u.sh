#!/bin/bash
# round 2, call U (2 GPUs): exactly what the driver runs for SCALE at N = 2 (both arms), plus the C4 workload tile-sharded
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29731 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_driver_ref_2gpu.log 2>&1; echo "ref exit $?"; tail -c 700 gpurun_out/r02_driver_ref_2gpu.log
timeout 900 $TR --nproc-per-node 2 --master-port 29732 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_driver_ours_2gpu.log 2>&1; echo "ours exit $?"; tail -c 1500 gpurun_out/r02_driver_ours_2gpu.log
timeout 900 $TR --nproc-per-node 2 --master-port 29733 bench.py --gpus 2 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-weak > gpurun_out/r02_bench_c4_2gpu_tiles.log 2>&1; echo "c4 exit $?"; tail -c 600 gpurun_out/r02_bench_c4_2gpu_tiles.log
timeout 600 python bench.py --workload c3-192 --batch 1 --steps 50 --no-cpu-baseline > gpurun_out/r02_bench_c3-192-pop1.log 2>&1; tail -c 700 gpurun_out/r02_bench_c3-192-pop1.log
