#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"k_lines" -c 1 -f -o gpurun_out/r1_klines2 python tools/shard_probe.py 1 300000 > gpurun_out/ncu_far.log 2>&1
tail -2 gpurun_out/ncu_far.log
