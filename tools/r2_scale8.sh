#!/bin/bash
# 8-GPU checks: depth-sharded flagship step (and the nu-sharded variant for comparison), 4 GPUs, grid sweep on 8 GPUs.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
A="--steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3"
( time timeout 600 $TR --nproc-per-node 8 --master-port 29601 bench.py --gpus 8 $A ) > gpurun_out/r2s_bench_8gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2s_bench_8gpu.log
( time timeout 600 $TR --nproc-per-node 8 --master-port 29602 bench.py --gpus 8 $A --partition nu ) > gpurun_out/r2s_bench_8gpu_nu.log 2>&1
python tools/bench_summary.py gpurun_out/r2s_bench_8gpu_nu.log
( time timeout 600 $TR --nproc-per-node 4 --master-port 29603 bench.py --gpus 4 $A ) > gpurun_out/r2s_bench_4gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2s_bench_4gpu.log
( time timeout 900 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --workload grid_sweep64 --steps 2 --warmup 1 ) > gpurun_out/r2s_sweep_8gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2s_sweep_8gpu.log
