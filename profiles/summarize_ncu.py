"""Key metrics of one kernel from an `ncu --set full` report.

    python profiles/summarize_ncu.py profiles/<report>.ncu-rep
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__inst_executed.sum.per_cycle_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2]
    print(f"report: {path}\nkernel: {data[hdr.index('Kernel Name')]}\n")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:66s} {data[i]:>18s} {units[i]}")
    print("\nstall reasons (warps per issue-active cycle):")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try:
                v = float(data[i])
            except ValueError:
                continue
            if v > 0.05:
                print(f"  {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
