#!/bin/bash
# First GPU call of the next round (everything here was written after round 1's GPU budget was spent).
#   gpurun --timeout 600 -- 'bash tools/gpu_next_round_first.sh'            (1 GPU part)
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/gpu_next_round_first.sh ddp'  (2-GPU DDP training step)
mkdir -p gpurun_out
if [ "$1" = "ddp" ]; then
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --workload train --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err
  echo "train N=2: exit $?"; cat gpurun_out/bench_train_n2.json; tail -3 gpurun_out/bench_train_n2.err
  exit 0
fi
# 1. the never-run tensor-core linear backward (csrc/backward_tc.cu)
CSE_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_backward_tc_gpu.py -q --tb=short -p no:cacheprovider \
    > gpurun_out/t_bwd_tc.log 2>&1
echo "backward_tc: exit $?"; tail -15 gpurun_out/t_bwd_tc.log
# 2. the whole GPU suite + smoke (training step included) as the driver runs them
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "pytest -m gpu: exit $?"; tail -3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke: exit $?"; tail -3 gpurun_out/smoke.log
# 3. training step (SURVEY §8d cfg 3): fixed 4 s with the ContExt loss, then the PIT variant and the 16 s max_sp_len cap
timeout 200 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err
echo "train N=1: exit $?"; cat gpurun_out/bench_train_n1.json
timeout 200 python bench.py --workload train --train-loss pit --steps 5 --warmup 3 > gpurun_out/bench_train_pit_n1.json 2>> gpurun_out/bench_train_n1.err
echo "train PIT: exit $?"; cat gpurun_out/bench_train_pit_n1.json
timeout 300 python bench.py --workload train --train-seconds 16 --steps 3 --warmup 3 > gpurun_out/bench_train_16s_n1.json 2>> gpurun_out/bench_train_n1.err
echo "train 16 s: exit $?"; cat gpurun_out/bench_train_16s_n1.json
# NOTE: an ncu launch list of the training bench costs ~0.2 s per launch (2000 launches per step): capture ONE step,
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 2100 --csv --log-file gpurun_out/launches_train.csv \
#       python bench.py --workload train --steps 1 --warmup 3
# and kill it after the first step (round 1 lost 7 GPU-minutes learning this).
# 4. A/B of the experimental tensor-core layers inside the whole training step (only if step 1 passed):
#   CSE_TRAIN_BF16=1 python bench.py --workload train --steps 5 --warmup 3
#   CSE_TRAIN_BF16=1 CSE_EXPERIMENTAL=1 python -m pytest tests/test_training_gpu.py -k whole_model -s   (expect the fp32
#   bounds to FAIL by design: compare the printed global rel-L2 with the reference's autocast drift 0.08-0.15)
