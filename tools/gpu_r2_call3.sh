#!/bin/bash
# Round 2, call 3: cost of overwriting vs accumulating tcgen05.mma, attention v3 tests + A/B timing + trace, fused optimiser.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. UMMA overwrite cost"
timeout 120 tools/micro/umma_pv_rate > $O/umma_pv_rate2.txt 2>&1; echo "umma_pv_rate: exit $?"; cat $O/umma_pv_rate2.txt
echo "== 2. attention v3 unit tests"
timeout 600 python -m pytest tests/test_stages_gpu.py -q -x --tb=short -p no:cacheprovider -k "test_attention and tc3" > $O/t_attn_v3.log 2>&1
echo "attention v3 tests: exit $?"; tail -5 $O/t_attn_v3.log
echo "== 3. attention A/B timing + trace"
timeout 300 python tools/attn_trace.py > $O/attn_trace3.txt 2>&1; echo "attn_trace: exit $?"; head -12 $O/attn_trace3.txt; tail -4 $O/attn_trace3.txt
ATTN_TRACE_MODE=3 timeout 300 python tools/attn_trace.py 2>&1 | tail -8 > $O/attn_trace3_mode3.txt; tail -4 $O/attn_trace3_mode3.txt
echo "== 4. fused optimiser"
timeout 300 python -m pytest tests/test_optim.py -q --tb=short -p no:cacheprovider > $O/t_optim.log 2>&1; echo "optim: exit $?"; tail -8 $O/t_optim.log
echo "== 5. bench (v3 + zero S default / v3 / v1)"
for v in 4 3 1; do
  CSE_ATTN_VER=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-train > $O/bench_attn_ver$v.json 2> $O/bench_attn_ver$v.err; echo "bench ver $v: exit $?"; cut -c1-120 $O/bench_attn_ver$v.json
done
