#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_training_gpu.py tests/test_optim.py -m gpu -x -q -k "graphed or optim" > $O/graph_test.log 2>&1
echo "test exit $?"; tail -15 $O/graph_test.log
for g in auto off; do
  timeout 300 python bench.py --workload train --steps 10 --warmup 3 --train-graph $g > $O/train_graph_$g.json 2> $O/train_graph_$g.err
  echo "train-graph $g: exit $?"; tail -2 $O/train_graph_$g.err
  python -c "
import json
d = json.loads(open('$O/train_graph_$g.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 2), round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), d['step_execution'], d['loss'], d['clocks'])"
done
