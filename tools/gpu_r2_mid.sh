#!/bin/bash
# mid-round check: the GPU test suite + the default bench line
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > $O/t_all.log 2>&1; echo "pytest -m gpu: exit $?"; tail -6 $O/t_all.log | cut -c1-300
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench (default flags): exit $?"; tail -2 $O/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("ours", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["clocks"], "parity", d["parity"]["ok"], d["parity"]["bf16_rel_l2"], "launches", d["gpu_launches"])
t = d["train"]
print("train sub-object", round(t["ms_per_step"], 2), round(t["value"], 1), "e2e", round(t["e2e"]["value"], 1), t["clocks"], t["gpu_launches"])
print(json.dumps(d["roofline"]))
PY
