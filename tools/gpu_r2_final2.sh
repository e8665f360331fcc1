#!/bin/bash
# Round 2, final evidence pass (after the LayerNorm-in-FFN / tail-kernel / graphed-training-step changes):
# bench lines, ncu launch list of the bench command, ncu --set full rows of one intra + one inter layer, of the
# memory-bound forward kernels and of one layer's backward kernels.  ncu only after the same command exited 0 without it;
# reports are converted to CSV here and deleted (gpurun copies back at most 64 MiB).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== 1. bench (ours, reference arm), training legs"
timeout 600 python bench.py --steps 30 --warmup 3 > $O/r02_bench.json 2> $O/r02_bench.err; echo "bench: exit $?"; cut -c1-160 $O/r02_bench.json; tail -2 $O/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference arm: exit $?"; cut -c1-200 $O/r02_bench_reference.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > $O/r02_train_4s.json 2> $O/r02_train_4s.err; echo "train 4 s: exit $?"; cut -c1-170 $O/r02_train_4s.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-graph off > $O/r02_train_4s_eager.json 2> $O/r02_train_4s_eager.err; echo "train 4 s eager: exit $?"; cut -c1-170 $O/r02_train_4s_eager.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-ragged > $O/r02_train_ragged.json 2> $O/r02_train_ragged.err; echo "train ragged: exit $?"; cut -c1-170 $O/r02_train_ragged.json
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-loss pit > $O/r02_train_pit.json 2> $O/r02_train_pit.err; echo "train pit: exit $?"; cut -c1-170 $O/r02_train_pit.json
timeout 300 python bench.py --workload train --steps 5 --warmup 3 --train-seconds 16 > $O/r02_train_16s.json 2> $O/r02_train_16s.err; echo "train 16 s: exit $?"; cut -c1-170 $O/r02_train_16s.json
timeout 300 python tools/train_profile.py > $O/r02_train_kernels.txt 2>&1; echo "train kernel table: exit $?"
echo "== 2. ncu launch list of the bench command"
timeout 200 python bench.py --steps 2 --warmup 3 --no-train > $O/bench_short.json 2> $O/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-train > $O/ncu_bench.log 2>&1
echo "ncu launch list: exit $?"; wc -l $O/r02_launches_bench.csv
echo "== 3. ncu --set full: layer kernels (intra layer 1, inter layer 1 of the 3rd forward)"
R='regex:gemm_tc_kernel|ffn_tc_kernel|attention_tc_kernel|attention_bf16_kernel'
timeout 200 python tools/quick_time.py 16 32000 bf16 1 > $O/plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none -k "$R" -s 271 -c 4 -o $O/r02_intra -f python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_intra.log 2>&1
echo "ncu intra: exit $?"
ncu -i $O/r02_intra.ncu-rep --page raw --csv > $O/r02_intra.csv 2>/dev/null; rm -f $O/r02_intra.ncu-rep
timeout 600 ncu --set full --clock-control none -k "$R" -s 303 -c 4 -o $O/r02_inter -f python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_inter.log 2>&1
echo "ncu inter: exit $?"
ncu -i $O/r02_inter.ncu-rep --page raw --csv > $O/r02_inter.csv 2>/dev/null; rm -f $O/r02_inter.ncu-rep
echo "== 4. ncu --set full: memory-bound forward kernels (2nd forward)"
R='regex:encoder_kernel|gn_apply_kernel|segment_kernel|build_seq_kernel|context_map_kernel|finish_stats_kernel|finish_apply_kernel|pred_head_kernel|prelu_ola_kernel|gate_kernel|decode_frames|decode_ola_kernel'
timeout 900 ncu --set full --clock-control none -k "$R" -s 24 -c 24 -o $O/r02_membound -f python tools/quick_time.py 16 32000 bf16 1 > $O/ncu_membound.log 2>&1
echo "ncu memory-bound kernels: exit $?"
ncu -i $O/r02_membound.ncu-rep --page raw --csv > $O/r02_membound.csv 2>/dev/null; rm -f $O/r02_membound.ncu-rep
echo "== 5. ncu --set full: one layer's backward kernels (3rd eager training step)"
timeout 200 python tools/train_steps_probe.py 32000 > $O/probe.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 951 -c 11 -o $O/r02_bwd_gemm -f python tools/train_steps_probe.py 32000 > $O/ncu_bwd_gemm.log 2>&1
echo "ncu backward GEMMs: exit $?"
ncu -i $O/r02_bwd_gemm.ncu-rep --page raw --csv > $O/r02_bwd_gemm.csv 2>/dev/null; rm -f $O/r02_bwd_gemm.ncu-rep
timeout 900 ncu --set full --clock-control none -k 'regex:attention_bwd_mma_kernel|colsum_cast_kernel|layernorm_bwd_kernel|optim_' -s 600 -c 14 -o $O/r02_bwd_other -f python tools/train_steps_probe.py 32000 > $O/ncu_bwd_other.log 2>&1
echo "ncu backward others: exit $?"
ncu -i $O/r02_bwd_other.ncu-rep --page raw --csv > $O/r02_bwd_other.csv 2>/dev/null; rm -f $O/r02_bwd_other.ncu-rep
ls -la $O/*.csv
