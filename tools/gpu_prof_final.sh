#!/bin/bash
# ncu --set full captures of one intra layer and one inter layer (3rd forward of the process), only after the
# same command exited 0 without ncu.  Matching launches per forward: 69 gemm_tc + 32 ffn_tc + 32 attention + 64 LN = 197.
mkdir -p gpurun_out
R='regex:gemm_tc_kernel|ffn_tc_kernel|attention_tc_kernel|attention_bf16_kernel|layernorm_kernel'
python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/plain.log 2>&1 || exit 1
cat gpurun_out/plain.log
ncu --set full --clock-control none --import-source on -k "$R" -s 401 -c 6 -o gpurun_out/r01_intra -f \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu_intra.log 2>&1
echo "ncu intra: exit $?"
ncu --set full --clock-control none --import-source on -k "$R" -s 449 -c 6 -o gpurun_out/r01_inter -f \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu_inter.log 2>&1
echo "ncu inter: exit $?"
ls -la gpurun_out/*.ncu-rep
