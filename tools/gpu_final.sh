#!/bin/bash
# Round-end validation on the GPU box: full GPU test suite, smoke(), both bench arms, then the ncu launch
# list of the same bench command (only after it exited 0 without ncu).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "pytest -m gpu: exit $?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke: exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench: exit $?"; cat gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref: exit $?"; cat gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list: exit $?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt
