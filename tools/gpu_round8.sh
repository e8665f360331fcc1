#!/bin/bash
mkdir -p gpurun_out
for t in stages forward; do
  timeout 900 python -m pytest tests/test_${t}_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_${t}.log 2>&1
  echo "${t}: exit $?"; grep -vE "mbarrier timeout" gpurun_out/t_${t}.log | grep -E "passed|failed|FAILED" | head
done
python tools/quick_time.py 16 32000 bf16 5 graph 2>&1 | tail -1
python tools/quick_time.py 1 32000 bf16 20 graph 2>&1 | tail -1
python tools/quick_time.py 1 32000 bf16 20 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench: exit $?"; tail -2 gpurun_out/bench.log | cut -c1-1500
