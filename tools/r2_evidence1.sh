#!/bin/bash
# Round-2 evidence on ONE B200: GPU tests, default bench, reference arm, every workload, grid sweep, sanitizer, ncu.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2e_gpu_tests.log 2>&1; tail -4 gpurun_out/r2e_gpu_tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2e_bench_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r2e_bench_1gpu.log
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2e_bench_reference.log 2>&1; tail -2 gpurun_out/r2e_bench_reference.log | cut -c1-400
for wl in sim10aa sim100aa solar_weak astar coolgiant_ir; do
  ( time timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --cpu-seconds 8 ) > gpurun_out/r2e_bench_$wl.log 2>&1
  python tools/bench_summary.py gpurun_out/r2e_bench_$wl.log | head -3
done
( time timeout 900 python bench.py --workload grid_sweep64 --steps 1 --warmup 1 ) > gpurun_out/r2e_sweep_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r2e_sweep_1gpu.log | head -3
( time timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2e_sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/r2e_sanitizer_memcheck_smoke.log
( time timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2e_sanitizer_racecheck_smoke.log 2>&1; tail -3 gpurun_out/r2e_sanitizer_racecheck_smoke.log
( time timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -q -k "alan or far" 2>&1 | tail -12 ) > gpurun_out/r2e_sanitizer_racecheck_k2_tests.log 2>&1; tail -4 gpurun_out/r2e_sanitizer_racecheck_k2_tests.log
bash tools/r2_profile.sh r2e
