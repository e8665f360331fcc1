#!/bin/bash
# tail-kernel changes (gate fast math, decode staging, finish two rows per trip): parity tests + kernel table
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_forward_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -x -q > $O/tail_tests.log 2>&1
echo "tests exit $?"; tail -3 $O/tail_tests.log
timeout 300 python tools/forward_kernels.py > $O/tail_kernels.txt 2>&1
echo "kernels exit $?"; head -30 $O/tail_kernels.txt
