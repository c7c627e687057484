"""Top stall sites of a kernel from an `ncu --set full --import-source on` report (SASS page):

    python tools/ncu_hot.py <report.ncu-rep> [top_n]

Prints the SASS instructions with the most warp-stall samples and the dominant stall reason of each."""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, ln in enumerate(out) if ln.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(out)) if out[i].startswith('"Kernel Name"')), len(out))
    rows = list(csv.DictReader(out[start:end]))
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    total = sum(int(r["# Samples"] or 0) for r in rows)
    print(f"{len(rows)} SASS instructions, {total} samples")
    agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
    print("by reason:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    idx = {id(r): i for i, r in enumerate(rows)}
    for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        n = int(r["# Samples"] or 0)
        why = max(stall_cols, key=lambda c: int(r[c] or 0))
        print(f"{idx[id(r)]:6d} {n:6d} {100 * n / total:5.1f}%  {why[6:]:14s} {r['Source'][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
