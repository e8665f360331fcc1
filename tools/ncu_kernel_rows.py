"""One line per kernel launch from an `ncu --set full ... --page raw --csv` dump: duration, DRAM bytes and GB/s,
tensor-pipe / XU utilisation, registers — the roofline evidence rows of profiles/r02_ncu_kernels.txt.

    ncu -i x.ncu-rep --page raw --csv > /tmp/x.csv ; python tools/ncu_kernel_rows.py /tmp/x.csv "title" [PEAK_GBS]
"""
import csv
import sys


def main():
    path, title = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6460.2
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale=1.0):
        try:
            v = float(r[col[name]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
        u = units[col[name]]
        if u.startswith("Gbyte"):
            v *= 1e9
        elif u.startswith("Mbyte"):
            v *= 1e6
        elif u.startswith("Kbyte"):
            v *= 1e3
        elif u.startswith("ms"):
            v *= 1e3
        elif u.startswith("ns"):
            v *= 1e-3
        return v * scale

    print(f"== {title}")
    print(f"{'kernel':58s} {'grid':>10s} {'regs':>5s} {'us':>8s} {'dram MB':>9s} {'GB/s':>7s} {'of HBM':>7s} {'tensor%':>8s} {'XU%':>6s} {'issue%':>7s}")
    for r in data:
        name = r[col["Kernel Name"]].replace("void ", "").replace("cse::", "").replace("(anonymous namespace)::", "")
        name = name.split("(")[0][:58]
        us = val(r, "gpu__time_duration.sum")
        by = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        gbs = by / (us * 1e-6) / 1e9
        grid = r[col["Grid Size"]].replace(" ", "")
        print(f"{name:58s} {grid:>10s} {val(r, 'launch__registers_per_thread'):5.0f} {us:8.1f} {by / 1e6:9.1f} {gbs:7.0f} "
              f"{gbs / peak:7.2f} {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} "
              f"{val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f}")
    print()


if __name__ == "__main__":
    main()
