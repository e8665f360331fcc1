"""Tabulate an `ncu --set full` report of consecutive kernel launches (one transformer layer) as columns.

    ncu -i gpurun_out/r02_intra.ncu-rep --page raw --csv > /tmp/intra.csv
    python tools/ncu_layer_table.py /tmp/intra.csv "intra layer (n = 251 tokens per sequence)"
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
]


def main():
    path, title = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    w = 24
    print(f"== {title}")

    def short(name):
        name = name.replace("void ", "").replace("cse::", "").replace("(anonymous namespace)::", "")
        return name.split("(")[0][:w]

    print(f"{'Kernel Name':78s}" + " |".join(f"{short(r[col['Kernel Name']]):>{w}s}" for r in data))
    for key in ("Grid Size", "Block Size"):
        print(f"{key:78s}" + " |".join(f"{r[col[key]]:>{w}s}" for r in data))
    for m in METRICS:
        if m not in col:
            continue
        vals = []
        for r in data:
            try:
                vals.append(f"{float(r[col[m]].replace(',', '')):.1f}")
            except ValueError:
                vals.append(r[col[m]])
        print(f"{m:63s}{units[col[m]][:14]:15s}" + " |".join(f"{v:>{w}s}" for v in vals))
    print()


if __name__ == "__main__":
    main()
