#!/bin/bash
# flagship step on N GPUs (depth partition), bounded C-port parity sample
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node $N --master-port $((29730+N)) bench.py --gpus $N --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3 ) > gpurun_out/r3s_bench_${N}gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r3s_bench_${N}gpu.log 2>&1 | sed -n '1,3p;7p' | cut -c1-420
