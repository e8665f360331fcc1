#!/bin/bash
# N = 2, the driver's command line: headline forward line with the `train` sub-object (captured DDP step, run last under a watchdog)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 30 --warmup 3 > $O/r02_bench_n2.json 2> $O/r02_bench_n2.err
echo "bench N=2: exit $?"; tail -3 $O/r02_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print("forward", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["clocks"], "parity", d["parity"]["ok"])
t = d["train"]
print("train", {k: t[k] for k in ("ms_per_step", "value", "step_execution", "comm") if k in t} if "error" not in t else t)
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 2 --workload train --steps 10 --warmup 3 > $O/r02_train_n2_eager.json 2> $O/r02_train_n2_eager.err
echo "train N=2 eager: exit $?"; python -c "
import json
d = json.loads(open('gpurun_out/r02_train_n2_eager.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 2), round(d['value'], 1), d['step_execution'], d['comm'])"
