"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches_train.csv STEPS > profiles/....txt

Prints launches / total time / share per kernel, per step (the capture spans STEPS identical steps).
ncu times are cold-cache and serialised: the SHARES are what carries over to an un-profiled run.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {len(rows)} launches over {steps} steps, {total / steps / 1e3:.2f} ms of kernel time per step (ncu, serialised)")
    print(f"# {'kernel':70s} {'launches/step':>13s} {'us/step':>10s} {'us/launch':>10s} {'share':>7s}")
    for name in sorted(tot, key=tot.get, reverse=True):
        print(f"{name[:70]:72s} {cnt[name] / steps:13.1f} {tot[name] / steps:10.1f} {tot[name] / cnt[name]:10.1f} "
              f"{100 * tot[name] / total:6.1f}%")


if __name__ == "__main__":
    main()
