#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -x -q -k "ffn" > $O/ffnln_unit.log 2>&1
echo "unit exit $?"; tail -3 $O/ffnln_unit.log
for v in 1 0 1; do
  CSE_FFN_LN=$v timeout 300 python tools/quick_time.py 16 32000 bf16 10 graph > $O/ffnln_time_$v.log 2>&1
  echo "CSE_FFN_LN=$v: exit $?"; tail -1 $O/ffnln_time_$v.log
done
CSE_FFN_LN=1 timeout 300 python tools/forward_kernels.py > $O/ffnln_kernels.txt 2>&1; sed -n 3,9p $O/ffnln_kernels.txt
