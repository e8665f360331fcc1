#!/bin/bash
# ncu --set full of (a) the memory-bound kernels of one cfg 2 forward, (b) the kernels of one autocast training step's
# layer backward, (c) the loop-side kernels.  Each after the same command exited 0 without ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/quick_time.py 16 32000 bf16 1 > $O/plain2.log 2>&1 || exit 1
R='regex:encoder_kernel|gn_apply_kernel|segment_kernel|build_seq_kernel|context_map_kernel|finish_stats_kernel|finish_apply_kernel|pred_head_kernel|prelu_ola_kernel|gate_kernel|decode_frames|decode_ola_kernel'
timeout 900 ncu --set full --clock-control none -k "$R" -s 36 -c 24 -o $O/r02_membound -f python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_membound.log 2>&1
echo "ncu memory-bound kernels: exit $?"
ncu -i $O/r02_membound.ncu-rep --page raw --csv > $O/r02_membound.csv 2>/dev/null; rm -f $O/r02_membound.ncu-rep
timeout 200 python tools/train_steps_probe.py 32000 > $O/probe.log 2>&1 || exit 1
R2='regex:gemm_tc_kernel|attention_bwd_mma_kernel|colsum_cast_kernel|layernorm_bwd_kernel|optim_'
timeout 900 ncu --set full --clock-control none -k "$R2" -s 2600 -c 30 -o $O/r02_trainbwd -f python tools/train_steps_probe.py 32000 > $O/ncu_trainbwd.log 2>&1
echo "ncu training backward kernels: exit $?"
ncu -i $O/r02_trainbwd.ncu-rep --page raw --csv > $O/r02_trainbwd.csv 2>/dev/null; rm -f $O/r02_trainbwd.ncu-rep
ls -la $O/r02_membound.csv $O/r02_trainbwd.csv   # (the reports themselves exceed what gpurun copies back)
