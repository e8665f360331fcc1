#!/bin/bash
mkdir -p gpurun_out
for t in stages forward; do
  timeout 900 python -m pytest tests/test_${t}_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/t_${t}.log 2>&1
  echo "${t}: exit $?"; tail -n 3 gpurun_out/t_${t}.log
done
grep "^\[bf16" gpurun_out/t_forward.log
python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu.log 2>&1
echo "ncu: exit $?"; cat gpurun_out/plain.log; tail -2 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
