#!/bin/bash
# ncu launch list (per-kernel durations) of one flagship step
TAG=${1:-r3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-direct > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_launch.log | cut -c1-300
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r) > 10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
out = []
for r in rows:
    out.append((r[4][:60], r[8], r[7], r[-1]))
for o in out[-80:]:
    print(o)
PY
