#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"k_far_coeffs" -c 3 -f -o gpurun_out/r1_far3 python tools/shard_probe.py 1 300000 > gpurun_out/ncu_far.log 2>&1
tail -2 gpurun_out/ncu_far.log
