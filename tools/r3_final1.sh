#!/bin/bash
# final state of the session on ONE B200: full GPU suite, smoke, default bench (direct mode + numba leg), launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r3z_gpu_tests.log 2>&1; grep -E "passed|failed" gpurun_out/r3z_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r3z_bench_1gpu.log 2>&1; python tools/bench_summary.py gpurun_out/r3z_bench_1gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3z_launches_bench_1gpu.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-direct > gpurun_out/r3z_ncu_launch.log 2>&1
tail -1 gpurun_out/r3z_ncu_launch.log | cut -c1-120
