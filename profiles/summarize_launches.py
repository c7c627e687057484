"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python profiles/summarize_launches.py profiles/<launches>.csv
"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:64]
        ns = float(row["Metric Value"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    print(f"{'kernel':66s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:66s} {n:5d} {ns / 1e3:10.1f} {ns / n / 1e3:9.1f} {100 * ns / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
