#!/bin/bash
# Round 2: fused out-proj + residual + LayerNorm epilogue.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. unit test"
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q --tb=short -p no:cacheprovider -x -k "residual_ln" > $O/t_resln.log 2>&1
echo "exit $?"; tail -8 $O/t_resln.log
echo "== 2. forward parity with the fused layer"
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_baseline_shapes_gpu.py -q --tb=short -p no:cacheprovider -x > $O/t_fwd_fused.log 2>&1
echo "exit $?"; tail -5 $O/t_fwd_fused.log
echo "== 3. bench A/B"
timeout 300 python bench.py --steps 10 --warmup 3 --no-train > $O/bench_resln.json 2> $O/bench_resln.err; echo "fused: exit $?"; cut -c1-120 $O/bench_resln.json
CSE_OUTPROJ_LN=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-train > $O/bench_resln_off.json 2> $O/bench_resln_off.err; echo "unfused: exit $?"; cut -c1-120 $O/bench_resln_off.json
python - <<'PY'
import json
for f in ("bench_resln", "bench_resln_off"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f, d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], r["frac"], r["classes"], d["parity"]["ok"], d["parity"]["bf16_rel_l2"])
PY
