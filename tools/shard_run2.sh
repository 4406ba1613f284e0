#!/bin/bash
SD_FAR_SCAN_ONLY=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/shard1_scan.csv python tools/shard_probe.py 1 300000 > /dev/null 2>&1
