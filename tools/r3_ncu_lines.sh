#!/bin/bash
# ncu --set full of ONE k_lines launch of the flagship step + per-source-line table
TAG=${1:-r3}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_lines" -s 1 -c 1 -f -o gpurun_out/${TAG}_lines \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_lines.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_lines.log | cut -c1-200
python tools/ncu_lines_agg.py gpurun_out/${TAG}_lines.ncu-rep k_lines 60 > gpurun_out/${TAG}_lines_src.txt 2>&1
ncu -i gpurun_out/${TAG}_lines.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
want=('Duration','Registers Per Thread','Achieved Occupancy','Theoretical Occupancy','Executed Ipc Active','Issue Slots Busy','Avg. Active Threads Per Warp','Avg. Not Predicated Off Threads Per Warp','Local Load','Local Store','L1/TEX Hit Rate','Shared Memory Configuration Size','Dynamic Shared Memory Per Block','Block Limit Shared Mem','Block Limit Registers','Warp Cycles Per Issued Instruction','No Eligible','One or More Eligible','FP64')
for r in csv.reader(sys.stdin):
    if len(r)>14 and any(w in r[12] for w in want): print(r[12],'=',r[14],r[13])
" > gpurun_out/${TAG}_lines_details.txt
cat gpurun_out/${TAG}_lines_details.txt
head -64 gpurun_out/${TAG}_lines_src.txt
