#!/bin/bash
# experiment batch on 2 GPUs: tests, 1-GPU bench (default and nsplit x2), 2-GPU bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/r2x_gpu_tests.log 2>&1; tail -4 gpurun_out/r2x_gpu_tests.log
A="--steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3"
( timeout 600 python bench.py $A ) > gpurun_out/r2x_bench_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r2x_bench_1gpu.log | grep -v "roofline_hbm\|clocks"
( SD_FAR_NSPLIT_SCALE=2 timeout 600 python bench.py $A ) > gpurun_out/r2x_bench_1gpu_nsplit2.log 2>&1; python tools/bench_summary.py gpurun_out/r2x_bench_1gpu_nsplit2.log | grep "==\|kernel_ms\|parity"
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 $A ) > gpurun_out/r2x_bench_2gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r2x_bench_2gpu.log | grep "==\|kernel_ms\|phase_ms\|parity"
