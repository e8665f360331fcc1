#!/bin/bash
# Full verification pass: the GPU test suite, smoke(), the bench line (ours + reference arm), the training leg.
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > $O/t_all.log 2>&1; echo "pytest -m gpu: exit $?"; tail -6 $O/t_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke: exit $?"; tail -3 $O/smoke.log | cut -c1-300
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench (default flags): exit $?"; tail -2 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "reference arm: exit $?"
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > $O/train_final.json 2> $O/train_final.err; echo "train: exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("ours", round(d["ms_per_step"], 3), round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), d["clocks"], "parity", d["parity"]["ok"], d["parity"]["bf16_rel_l2"], "launches", d["gpu_launches"])
t = d["train"]
print("train sub-object", round(t["ms_per_step"], 2), round(t["value"], 1), "e2e", round(t["e2e"]["value"], 1), t["clocks"], t["gpu_launches"])
print("cpu_baseline", d["cpu_baseline"])
r = json.loads(open("gpurun_out/bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
t = json.loads(open("gpurun_out/train_final.json").read().strip().splitlines()[-1])
print("train standalone", round(t["ms_per_step"], 2), round(t["value"], 1), "e2e", round(t["e2e"]["value"], 1), t["clocks"], round(t["roofline"]["frac"], 3))
PY
