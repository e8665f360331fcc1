#!/bin/bash
# Round 2: tensor-core attention backward.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. tests"
timeout 600 python -m pytest tests/test_backward_tc_gpu.py -q --tb=short -p no:cacheprovider -x > $O/t_bwd_tc3.log 2>&1
echo "exit $?"; tail -12 $O/t_bwd_tc3.log | cut -c1-300
timeout 600 python -m pytest tests/test_training_gpu.py -q -s --tb=short -p no:cacheprovider -k "autocast" > $O/t_autocast_train3.log 2>&1
echo "autocast training: exit $?"; grep -E "^\[autocast|passed|failed|Error|assert" $O/t_autocast_train3.log | cut -c1-300 | tail -8
echo "== 2. train bench"
timeout 300 python bench.py --workload train --steps 5 --warmup 3 > $O/train_amp3.json 2> $O/train_amp3.err; echo "train: exit $?"; cut -c1-200 $O/train_amp3.json; tail -3 $O/train_amp3.err
echo "== 3. kernel table"
timeout 300 python tools/train_profile.py > $O/train_kernels3.txt 2>&1; echo "exit $?"; head -24 $O/train_kernels3.txt | cut -c1-200
