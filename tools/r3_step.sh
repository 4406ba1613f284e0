#!/bin/bash
bash tools/r3_fast.sh
bash tools/r3_launches.sh ${1:-r3d} | grep -v "cub::\|k_cls\|k_tile\|k_dnu\|k_line_pre\|k_depth_pre" | tail -32
