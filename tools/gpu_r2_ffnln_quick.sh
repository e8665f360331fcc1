#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for v in 1 2 0 1 2; do
  CSE_FFN_LN=$v timeout 300 python tools/quick_time.py 16 32000 bf16 10 graph > $O/ffnln_time_$v.log 2>&1
  echo "CSE_FFN_LN=$v: exit $?"; tail -1 $O/ffnln_time_$v.log
done
