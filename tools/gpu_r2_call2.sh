#!/bin/bash
# Round 2, call 2: autocast training tests (fixture loader fixed), UMMA shape microbenchmark for the attention PV
# step, ncu launch list of the autocast training step.
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. autocast training tests"
timeout 600 python -m pytest tests/test_training_gpu.py -q -s --tb=short -p no:cacheprovider -k "autocast" > $O/t_autocast_train.log 2>&1
echo "autocast training: exit $?"; grep -E "^\[autocast|passed|failed|Error|assert" $O/t_autocast_train.log | cut -c1-400 | tail -20
echo "== 2. UMMA shapes"
timeout 120 tools/micro/umma_pv_rate > $O/umma_pv_rate.txt 2>&1; echo "umma_pv_rate: exit $?"; cat $O/umma_pv_rate.txt
echo "== 3. training launch list"
bash tools/gpu_train_prof.sh
