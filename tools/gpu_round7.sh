#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stages_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k attention > gpurun_out/t_attn.log 2>&1
echo "attn: exit $?"; grep -vE "mbarrier timeout" gpurun_out/t_attn.log | grep -E "passed|failed|FAILED|assert " | head -20; grep -c "mbarrier timeout" gpurun_out/t_attn.log
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
for n,v,g in names[idx+10:idx+18]: print(f"{v:9.1f} us  grid {g:16s} {n[:60]}")
k=[i for i in range(idx,len(names)) if "finish_apply" in names[i][0]][0]
for n,v,g in names[k+1:k+8]: print(f"{v:9.1f} us  grid {g:16s} {n[:60]}")
PY
