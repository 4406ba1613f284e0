#!/bin/bash
# quick validation on ONE B200: far-field stress tests first, then the GPU suite, then a short flagship bench
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_farfield_stress.py -m gpu -q 2>&1 | tail -120 ) > gpurun_out/r3q_stress.log 2>&1
tail -40 gpurun_out/r3q_stress.log
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -120 ) > gpurun_out/r3q_gpu_tests.log 2>&1
tail -30 gpurun_out/r3q_gpu_tests.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r3q_bench_1gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r3q_bench_1gpu.log
( time timeout 600 python bench.py --workload solar_weak --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r3q_bench_weak.log 2>&1
python tools/bench_summary.py gpurun_out/r3q_bench_weak.log
