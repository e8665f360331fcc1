#!/bin/bash
# Round 2, call 4: attention v4 tests + A/B timing + trace; bench per version.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. attention v4 unit tests"
timeout 600 python -m pytest tests/test_stages_gpu.py -q -x --tb=short -p no:cacheprovider -k "test_attention and tc4" > $O/t_attn_v4.log 2>&1
echo "attention v4 tests: exit $?"; tail -5 $O/t_attn_v4.log
echo "== 2. attention A/B timing + trace"
timeout 300 python tools/attn_trace.py > $O/attn_trace4.txt 2>&1; echo "attn_trace: exit $?"; head -12 $O/attn_trace4.txt; tail -5 $O/attn_trace4.txt
ATTN_TRACE_MODE=6 timeout 300 python tools/attn_trace.py 2>&1 | tail -6 > $O/attn_trace4_mode6.txt; tail -5 $O/attn_trace4_mode6.txt
echo "== 3. bench per attention version"
for v in 4 5 6 1; do
  CSE_ATTN_VER=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-train > $O/bench_attn_ver$v.json 2> $O/bench_attn_ver$v.err; echo "bench ver $v: exit $?"; cut -c1-120 $O/bench_attn_ver$v.json
done
