#!/bin/bash
# One gpurun call: every GPU test file in its own process (a trapped kernel poisons its context),
# logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for t in stages gemm_tc losses forward; do
  timeout 900 python -m pytest tests/test_${t}_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_${t}.log 2>&1
  echo "${t}: exit $?"
  tail -n 3 gpurun_out/t_${t}.log
done
timeout 600 python tools/quick_time.py > gpurun_out/time.log 2>&1; echo "time: exit $?"; cat gpurun_out/time.log
