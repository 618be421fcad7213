"""Summarise ncu outputs into profiles/: a per-kernel launch table from a `--metrics gpu__time_duration.sum`
CSV log, and the key counters of every launch in a `--set full` .ncu-rep report."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "lts__t_bytes.sum", "sm__cycles_active.avg",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[start + 1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "us" else v / 1e6 if r[ui] == "ns" else v
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1; a[1] += v; tot += v
    print(f"launches: {len(rows) - start - 1}, total device time {tot:.3f} ms (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {v:.3f} | {100 * v / tot:.1f}% |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for d in rows[2:]:
        print(f"## {d[ki][:110]}")
        for k in KEYS:
            if k in h:
                print(f"  {k:78s} {d[h.index(k)]:>16s} {u[h.index(k)]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2])
