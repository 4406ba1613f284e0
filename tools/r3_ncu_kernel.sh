#!/bin/bash
# ncu --set full of launches of one kernel (regex $2, skip $3, count $4) of the flagship step + per-source-line table
TAG=${1:-r3}; RX=${2:-k_lines}; SKIP=${3:-1}; CNT=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log | cut -c1-120
python tools/ncu_lines_agg.py gpurun_out/${TAG}.ncu-rep "$RX" 45 > gpurun_out/${TAG}_src.txt 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
want=('Duration','Registers Per Thread','Achieved Occupancy','Theoretical Occupancy','Executed Ipc Active','Issue Slots Busy','Avg. Active Threads Per Warp','Avg. Not Predicated Off Threads Per Warp','L1/TEX Hit Rate','Block Limit','Warp Cycles Per Issued Instruction','No Eligible','Grid Size','DRAM Throughput','Compute (SM) Throughput')
for r in csv.reader(sys.stdin):
    if len(r)>14 and any(w in r[12] for w in want): print(r[0], r[12],'=',r[14],r[13])
" > gpurun_out/${TAG}_details.txt
cat gpurun_out/${TAG}_details.txt
head -50 gpurun_out/${TAG}_src.txt
