#!/bin/bash
# Round-2 evidence on EIGHT B200s: flagship step at 2 / 4 / 8 GPUs (depth partition; nu partition at 8 for comparison), sweep at 8.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
A="--steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3"
P=29700
for n in 8 4 2; do
  P=$((P+1))
  ( time timeout 600 $TR --nproc-per-node $n --master-port $P bench.py --gpus $n $A ) > gpurun_out/r2e_bench_${n}gpu.log 2>&1
  python tools/bench_summary.py gpurun_out/r2e_bench_${n}gpu.log | head -4
done
( time timeout 600 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 $A --partition nu ) > gpurun_out/r2e_bench_8gpu_nu.log 2>&1
python tools/bench_summary.py gpurun_out/r2e_bench_8gpu_nu.log | head -4
( time timeout 900 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --workload grid_sweep64 --steps 2 --warmup 1 ) > gpurun_out/r2e_sweep_8gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2e_sweep_8gpu.log | head -3
( time timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -5 ) > gpurun_out/r2e_gpu_multi_test.log 2>&1; tail -3 gpurun_out/r2e_gpu_multi_test.log
