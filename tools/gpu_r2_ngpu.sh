#!/bin/bash
# The driver's command line at N GPUs: headline forward line + the `train` sub-object (captured DDP step, run last under a
# watchdog).  usage: gpurun --gpus N -- 'bash tools/gpu_r2_ngpu.sh N'   (profiles/r02_bench_n{2,4,8}_ddp.json)
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 30 --warmup 3 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err
echo "bench N=$N: exit $?"; tail -2 $O/r02_bench_n$N.err | cut -c1-200
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n$N.json").read().strip().splitlines()[-1])
print("forward", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"], "parity", d["parity"]["ok"])
t = d["train"]
print("train", {k: t[k] for k in ("ms_per_step", "value", "step_execution") if k in t} if "error" not in t else t)
PY
