#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -15 ) > gpurun_out/r3m_gpu_multi_test.log 2>&1; tail -5 gpurun_out/r3m_gpu_multi_test.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 2 --master-port 29721 bench.py --gpus 2 --steps 10 --warmup 3 --no-direct --cpu-kind port --cpu-seconds 3 ) > gpurun_out/r3m_bench_2gpu.log 2>&1
python tools/bench_summary.py gpurun_out/r3m_bench_2gpu.log 2>&1 | sed -n '1,3p;7p' | cut -c1-400
