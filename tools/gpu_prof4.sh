#!/bin/bash
mkdir -p gpurun_out
python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 9 -c 1 -o gpurun_out/prof_g -f \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu_g.log 2>&1
echo "ncu: exit $?"
