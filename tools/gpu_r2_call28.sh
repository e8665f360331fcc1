#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -q --tb=short -p no:cacheprovider -x -k "mask_decode" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_baseline_shapes_gpu.py -q --tb=short -p no:cacheprovider -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_df.json 2> $O/bench_df.err; echo "bench: exit $?"
CSE_DECODE_SIMT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-train > $O/bench_df_simt.json 2> $O/bench_df_simt.err; echo "bench simt: exit $?"
python - <<'PY'
import json
for f in ("bench_df", "bench_df_simt"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 3), round(d["value"], 1), round(d["e2e"]["value"], 1), d["parity"]["ok"], d["parity"]["bf16_rel_l2"], d["parity"]["bf16_dsisnr_db"], d["clocks"], (d.get("train") or {}).get("ms_per_step"), (d.get("train") or {}).get("clocks"))
PY
