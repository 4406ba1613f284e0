#!/bin/bash
# final build: full GPU suite, smoke, flagship bench line (no direct mode, C-port sample), launch list
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r3y_gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/r3y_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r3y_bench_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r3y_bench_1gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3y_launches_bench_1gpu.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-direct > gpurun_out/r3y_ncu_launch.log 2>&1
tail -1 gpurun_out/r3y_ncu_launch.log | cut -c1-100
