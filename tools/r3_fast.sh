#!/bin/bash
# stress + kernel tests, then a short flagship / weak-line bench
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_farfield_stress.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r3f_tests.log 2>&1
tail -5 gpurun_out/r3f_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3 ) > gpurun_out/r3f_bench_1gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r3f_bench_1gpu.log
( timeout 600 python bench.py --workload solar_weak --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3 ) > gpurun_out/r3f_bench_weak.log 2>&1
python tools/bench_summary.py gpurun_out/r3f_bench_weak.log
