#!/bin/bash
# N = 8, the driver's command line (does the captured DDP step hold beyond N = 2?)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 30 --warmup 3 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err
echo "bench N=8: exit $?"; tail -2 $O/r02_bench_n8.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("forward", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"], "parity", d["parity"]["ok"])
t = d["train"]
print("train", {k: t[k] for k in ("ms_per_step", "value", "step_execution") if k in t} if "error" not in t else t)
PY
