#!/bin/bash
# N = 2: what the DDP wrapper costs on the (host launch-bound) training step, and which of its knobs matters.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --workload train --steps 10 --warmup 3 "$@" > gpurun_out/ddp_$name.json 2> gpurun_out/ddp_$name.err; echo "$name: exit $?"; PORT=$((PORT+1)); }
PORT=29530
run default
run bucket_view --ddp-bucket-view
run one_bucket --ddp-bucket-mb 200
run view_one_static --ddp-bucket-view --ddp-bucket-mb 200 --ddp-static-graph
run torch_optim --train-optim torch
python - <<'PY'
import json
for n in ("default", "bucket_view", "one_bucket", "view_one_static", "torch_optim"):
    try:
        d = json.loads(open(f"gpurun_out/ddp_{n}.json").read().strip().splitlines()[-1])
        c = d["comm"]
        print(f"{n:18s} step {d['ms_per_step']:.2f} ms  no_sync {c['step_ms_without_allreduce']:.2f}  exposed {c['exposed_allreduce_ms']:.2f}  alone {c['allreduce_alone_ms']:.2f}")
    except Exception as e:
        print(n, "ERR", e)
PY
