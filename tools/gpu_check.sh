#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
echo "gemm: exit $?"; grep -vE "mbarrier timeout" gpurun_out/t_gemm.log | grep -E "passed|failed|FAILED|assert " | head; grep -c "mbarrier timeout" gpurun_out/t_gemm.log
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/t_forward.log 2>&1
echo "forward: exit $?"; grep -E "passed|failed|FAILED" gpurun_out/t_forward.log | head
python tools/quick_time.py 16 32000 bf16 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu.log 2>&1
echo "ncu: exit $?"; cat gpurun_out/plain.log
python - <<'PY'
import csv, re
lines=[l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
rows=list(csv.DictReader(lines))
names=[(re.sub(r"\(.*","",r["Kernel Name"]), float(r["Metric Value"].replace(",",""))/1e3, r["Grid Size"]) for r in rows]
idx=[i for i,n in enumerate(names) if "encoder_kernel" in n[0]][-1]
for n,v,g in names[idx+10:idx+16]: print(f"{v:9.1f} us  grid {g:16s} {n[:60]}")
k=[i for i in range(idx,len(names)) if "finish_apply" in names[i][0]][0]
for n,v,g in names[k+1:k+7]: print(f"{v:9.1f} us  grid {g:16s} {n[:60]}")
PY
python tools/quick_time.py 16 32000 bf16 5 graph 2>&1 | tail -1
