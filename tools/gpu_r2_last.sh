#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_forward_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > $O/last_default.log 2>&1; echo "default: exit $?"; tail -2 $O/last_default.log
CSE_FFN_LN=1 timeout 300 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "bf16 or graph" > $O/last_force.log 2>&1; echo "forced: exit $?"; tail -2 $O/last_force.log
timeout 120 python tools/quick_time.py 1 16000 bf16 20 graph 2>&1 | tail -1
timeout 120 python tools/quick_time.py 16 32000 bf16 10 graph 2>&1 | tail -1
