#!/bin/bash
# Round 2: split-K wgrad + fused optimiser in the training step; kernel-time table of the step.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. backward tc tests (split-K wgrad)"
CSE_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_backward_tc_gpu.py tests/test_gemm_tc_gpu.py -q --tb=short -p no:cacheprovider -x > $O/t_bwd_tc2.log 2>&1
echo "exit $?"; tail -5 $O/t_bwd_tc2.log
timeout 600 python -m pytest tests/test_training_gpu.py -q -s --tb=short -p no:cacheprovider -k "autocast" > $O/t_autocast_train2.log 2>&1
echo "autocast training: exit $?"; grep -E "^\[autocast|passed|failed|Error|assert" $O/t_autocast_train2.log | cut -c1-300 | tail -8
echo "== 2. train bench"
timeout 300 python bench.py --workload train --steps 5 --warmup 3 > $O/train_amp2.json 2> $O/train_amp2.err; echo "train fused: exit $?"; cut -c1-200 $O/train_amp2.json; tail -3 $O/train_amp2.err
timeout 300 python bench.py --workload train --steps 5 --warmup 3 --train-optim torch > $O/train_amp2_torchopt.json 2> $O/train_amp2_torchopt.err; echo "train torch optim: exit $?"; cut -c1-200 $O/train_amp2_torchopt.json
echo "== 3. kernel table"
timeout 300 python tools/train_profile.py > $O/train_kernels.txt 2>&1; echo "exit $?"; head -45 $O/train_kernels.txt
