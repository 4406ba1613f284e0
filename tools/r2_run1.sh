#!/bin/bash
# Round 2, first GPU call: full GPU test-suite, default bench (parity block, executed-work rooflines, numba CPU leg),
# the other workloads, compute-sanitizer on the smoke pass.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
nproc
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests.log 2>&1
tail -5 gpurun_out/r2_gpu_tests.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_default.log 2>&1
tail -3 gpurun_out/r2_bench_default.log | cut -c1-1500
for wl in sim10aa sim100aa solar_weak astar coolgiant_ir; do
  ( time python bench.py --workload $wl --steps 5 --warmup 3 --cpu-seconds 8 ) > gpurun_out/r2_bench_$wl.log 2>&1
  tail -4 gpurun_out/r2_bench_$wl.log | cut -c1-600
done
( time python bench.py --workload grid_sweep64 --steps 1 --warmup 1 ) > gpurun_out/r2_bench_sweep.log 2>&1
tail -4 gpurun_out/r2_bench_sweep.log | cut -c1-800
( time timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1
tail -6 gpurun_out/r2_sanitizer_memcheck_smoke.log
( time timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_sanitizer_racecheck_smoke.log 2>&1
tail -6 gpurun_out/r2_sanitizer_racecheck_smoke.log
