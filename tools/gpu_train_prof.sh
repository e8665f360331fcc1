#!/bin/bash
# ncu launch list of the training bench command (only after it exited 0 without ncu): 6 identical steps
# (3 warm-up, 1 counted, 1 timed, 1 end-to-end) of `bench.py --workload train --steps 1 --warmup 3`.
mkdir -p gpurun_out
timeout 200 python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/bench_train_short.json 2> gpurun_out/bench_train_short.err || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
echo "ncu launch list: exit $?"
python tools/summarize_launches.py gpurun_out/launches_train.csv 6 > gpurun_out/launches_train_summary.txt
head -30 gpurun_out/launches_train_summary.txt
