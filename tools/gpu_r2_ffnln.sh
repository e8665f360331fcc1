#!/bin/bash
# the feed-forward kernel with its LayerNorm warps: unit test, forward parity, timing with and without
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -x -q -k "ffn" > $O/ffnln_unit.log 2>&1
echo "unit exit $?"; tail -3 $O/ffnln_unit.log
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_baseline_shapes_gpu.py tests/test_stages_gpu.py tests/test_backward_tc_gpu.py -m gpu -x -q > $O/ffnln_fwd.log 2>&1
echo "forward exit $?"; tail -3 $O/ffnln_fwd.log
for v in 1 0 1; do
  CSE_FFN_LN=$v timeout 300 python tools/quick_time.py 16 32000 bf16 10 graph > $O/ffnln_time_$v.log 2>&1
  echo "CSE_FFN_LN=$v: exit $?"; tail -1 $O/ffnln_time_$v.log
done
timeout 300 python tools/ffn_ln_time.py 2>&1 | head -9
