#!/bin/bash
# Round evidence on ONE B200: default bench, reference arm, ncu launch list of the bench command, full ncu captures.
set -x
python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log > gpurun_out/r1_bench_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log > gpurun_out/r1_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_far_coeffs|k_raytrace|k_continuum|k_broadening|k_build_records" -c 8 -f -o gpurun_out/r1_step_full python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"k_lines" -s 3 -c 1 -f -o gpurun_out/r1_direct_full python tools/gpu_probe.py 300000 > gpurun_out/ncu_direct.log 2>&1
cut -c1-400 gpurun_out/r1_bench_1gpu.json
