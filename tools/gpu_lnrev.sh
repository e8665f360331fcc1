#!/bin/bash
mkdir -p gpurun_out
for REV in 0 1; do
export CSE_LN_REVERSE=$REV
echo "=== CSE_LN_REVERSE=$REV"
python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"layernorm|gemm_tc|ffn_tc|attention" -s 300 -c 12 --csv --log-file gpurun_out/rev.csv \
    python tools/quick_time.py 16 32000 bf16 1 > gpurun_out/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/rev.csv") if not l.startswith("=="))]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value")
cur={}
for r in rows[1:]:
    key=r[0]
    cur.setdefault(key,[r[ik].split("(")[0][-40:]]).append(r[iv])
for k,v in cur.items(): print(v)
PY
python tools/quick_time.py 16 32000 bf16 5 graph 2>&1 | tail -1
done
