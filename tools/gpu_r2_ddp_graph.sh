#!/bin/bash
# N = 2: the DistributedDataParallel training step captured as one CUDA graph per rank (opt-in), against the eager step
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --workload train --steps 10 --warmup 3 --train-graph ddp > $O/r02_train_n2_graph.json 2> $O/r02_train_n2_graph.err
echo "ddp graph: exit $?"; tail -1 $O/r02_train_n2_graph.json | cut -c1-400; tail -5 $O/r02_train_n2_graph.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_train_n2_graph.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"], 2), round(d["value"], 1), d["step_execution"], d["comm"])
except Exception as e:
    print("no result:", e)
PY
