#!/bin/bash
# Round-3 (second session of round 2) evidence on ONE B200: GPU tests, default bench, reference arm, every workload, grid
# sweep, sanitizer, ncu launch list + full captures.
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r3e_gpu_tests.log 2>&1; tail -4 gpurun_out/r3e_gpu_tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r3e_bench_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r3e_bench_1gpu.log
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r3e_bench_reference.log 2>&1; tail -2 gpurun_out/r3e_bench_reference.log | cut -c1-400
for wl in sim10aa sim100aa solar_weak astar coolgiant_ir; do
  ( time timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 --cpu-seconds 8 ) > gpurun_out/r3e_bench_$wl.log 2>&1
  python tools/bench_summary.py gpurun_out/r3e_bench_$wl.log | head -3
done
( time timeout 900 python bench.py --workload grid_sweep64 --steps 1 --warmup 1 ) > gpurun_out/r3e_sweep_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r3e_sweep_1gpu.log | head -3
( time timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r3e_sanitizer_memcheck_smoke.log 2>&1; tail -3 gpurun_out/r3e_sanitizer_memcheck_smoke.log
( time timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r3e_sanitizer_racecheck_smoke.log 2>&1; tail -3 gpurun_out/r3e_sanitizer_racecheck_smoke.log
( time timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -q -k "alan or far" 2>&1 | tail -12 ) > gpurun_out/r3e_sanitizer_racecheck_k2_tests.log 2>&1; tail -4 gpurun_out/r3e_sanitizer_racecheck_k2_tests.log
TAG=r3e
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_1gpu.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_launch.log | cut -c1-200
# one step = 43 launches of our kernels + the CUB sort; capture the last step's main kernels
ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_far_coeffs|k_s2m|k_m2m|k_m2l|k_build_records|k_raytrace|k_continuum|k_broadening" \
    -s 42 -c 21 -f -o gpurun_out/${TAG}_step_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
ls -la gpurun_out/${TAG}_step_full.ncu-rep
