#!/bin/bash
# ncu evidence of the flagship step on ONE B200: launch list (all kernels) + --set full captures of the main kernels.
# Usage: bash tools/r2_profile.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_1gpu.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_launch.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:"k_lines|k_far_coeffs|k_build_records|k_raytrace|k_continuum|k_broadening" \
    -s 24 -c 8 -f -o gpurun_out/${TAG}_step_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
ls -la gpurun_out/${TAG}_step_full.ncu-rep
