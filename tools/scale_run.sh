#!/bin/bash
# multi-GPU scaling of bench.py (run under gpurun --gpus 8): usage scale_run.sh "4 8"
for n in ${1:-1 2 4 8}; do
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/scale_$n.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/scale_$n.json
  fi
  python - <<PY
import json
d=json.loads(open("gpurun_out/scale_$n.json").read())
print($n, "gpus: ms/step", round(d["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "direct ms", round(d["farfield"]["ms_per_step_direct"],1), "phase", {k: round(v,2) for k,v in d["phase_ms"].items()})
PY
done
