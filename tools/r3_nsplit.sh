#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_farfield_stress.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r3n_tests.log 2>&1; tail -2 gpurun_out/r3n_tests.log
for sc in 1 2 4; do
  ( SD_FAR_NSPLIT_SCALE=$sc timeout 300 python bench.py --steps 10 --warmup 3 --no-direct --no-cpu-baseline ) > gpurun_out/r3n_nsplit$sc.log 2>&1
  echo "NSPLIT_SCALE=$sc"; python tools/bench_summary.py gpurun_out/r3n_nsplit$sc.log 2>/dev/null | sed -n '1,2p' | cut -c1-300
done
