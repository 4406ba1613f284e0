#!/bin/bash
mkdir -p gpurun_out
for mb in 10 12; do
  ( SD_K2_MINB=$mb timeout 300 python bench.py --steps 5 --warmup 2 --no-direct --no-cpu-baseline ) > gpurun_out/r3x_minb$mb.log 2>&1
  echo "MINB=$mb"; python tools/bench_summary.py gpurun_out/r3x_minb$mb.log 2>/dev/null | head -2
done
bash tools/r3_ncu_kernel.sh r3e_lines k_lines 1 1 > gpurun_out/r3e_lines_out.txt 2>&1
head -75 gpurun_out/r3e_lines_out.txt
