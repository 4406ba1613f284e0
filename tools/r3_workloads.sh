#!/bin/bash
# full GPU suite + one bench line per workload on ONE B200
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r3w_gpu_tests.log 2>&1
tail -4 gpurun_out/r3w_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in sim10aa sim100aa astar coolgiant_ir solar_weak; do
  ( timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3 ) > gpurun_out/r3w_bench_$w.log 2>&1
  python tools/bench_summary.py gpurun_out/r3w_bench_$w.log 2>&1 | sed -n '1,2p;7p' | cut -c1-330
done
