#!/bin/bash
# Round 2 evidence pass: final bench lines (ours, reference arm, train legs), ncu launch list of the bench command,
# ncu --set full of one intra layer's kernels.  ncu runs only after the same command exited 0 without it.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== 1. bench (ours, reference arm)"
timeout 600 python bench.py --steps 30 --warmup 3 > $O/r02_bench.json 2> $O/r02_bench.err; echo "bench: exit $?"; cut -c1-160 $O/r02_bench.json; tail -3 $O/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference arm: exit $?"; cut -c1-300 $O/r02_bench_reference.json
echo "== 2. training legs"
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > $O/r02_train_4s.json 2> $O/r02_train_4s.err; echo "train 4 s: exit $?"; cut -c1-170 $O/r02_train_4s.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-ragged > $O/r02_train_ragged.json 2> $O/r02_train_ragged.err; echo "train ragged: exit $?"; cut -c1-170 $O/r02_train_ragged.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-loss pit > $O/r02_train_pit.json 2> $O/r02_train_pit.err; echo "train pit: exit $?"; cut -c1-170 $O/r02_train_pit.json
timeout 300 python bench.py --workload train --steps 5 --warmup 3 --train-seconds 16 > $O/r02_train_16s.json 2> $O/r02_train_16s.err; echo "train 16 s: exit $?"; cut -c1-170 $O/r02_train_16s.json
echo "== 3. ncu launch list of the bench command"
timeout 200 python bench.py --steps 2 --warmup 3 --no-train > $O/bench_short.json 2> $O/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-train > $O/ncu_bench.log 2>&1
echo "ncu launch list: exit $?"; wc -l $O/r02_launches_bench.csv
echo "== 4. ncu --set full of one intra layer (3rd forward)"
R='regex:gemm_tc_kernel|ffn_tc_kernel|attention_tc_kernel|attention_bf16_kernel|layernorm_kernel'
timeout 200 python tools/quick_time.py 16 32000 bf16 1 > $O/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k "$R" -s 401 -c 6 -o $O/r02_intra -f \
    python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_intra.log 2>&1
echo "ncu intra: exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "$R" -s 449 -c 6 -o $O/r02_inter -f \
    python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_inter.log 2>&1
echo "ncu inter: exit $?"
ls -la $O/*.ncu-rep
