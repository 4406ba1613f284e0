#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/shard_probe.py 1 300000
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/shard1.csv python tools/shard_probe.py 1 300000 > /dev/null 2>&1
python tools/shard_probe.py 8 300000 -1 bal
