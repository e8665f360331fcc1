#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
echo "gemm: exit $?"; tail -n 4 gpurun_out/t_gemm.log
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -s --tb=line -p no:cacheprovider > gpurun_out/t_forward.log 2>&1
echo "forward: exit $?"; tail -n 4 gpurun_out/t_forward.log
python tools/quick_time.py 16 32000 bf16 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu.log 2>&1
echo "ncu: exit $?"; cat gpurun_out/plain.log
