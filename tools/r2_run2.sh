#!/bin/bash
# Round 2: validate the new K1 / preparation pass (tests), 1-GPU and 2-GPU step times.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -80 ) > gpurun_out/r2b_gpu_tests.log 2>&1
tail -6 gpurun_out/r2b_gpu_tests.log
( time python bench.py --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r2b_bench_1gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2b_bench_1gpu.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 4 ) > gpurun_out/r2b_bench_2gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r2b_bench_2gpu.log
fi
