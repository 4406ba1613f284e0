#!/bin/bash
# quick validation on ONE B200: GPU tests + short flagship bench (no direct mode, C-port parity sample only)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -80 ) > gpurun_out/r2q_gpu_tests.log 2>&1
tail -6 gpurun_out/r2q_gpu_tests.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r2q_bench_1gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2q_bench_1gpu.log
( time timeout 600 python bench.py --workload solar_weak --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r2q_bench_weak.log 2>&1
python tools/bench_summary.py gpurun_out/r2q_bench_weak.log
