#!/bin/bash
mkdir -p gpurun_out
python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 7 -c 2 -o gpurun_out/prof_attn_tc -f \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn: exit $?"
