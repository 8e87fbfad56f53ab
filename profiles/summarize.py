"""Turn ncu outputs brought back in gpurun_out/ into the markdown summaries kept in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_x.csv [skip_substr ...]   # one step's launch list
    python profiles/summarize.py ncu gpurun_out/prof_x.ncu-rep                           # key metrics of one capture
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path, skip):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    out = [(r[ki], r[gi], r[bi], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]
           if "qsae" in r[ki] and not any(s in r[ki] for s in skip)]
    # keep the last full repetition: cut at the last occurrence of the first kernel name
    first = out[0][0]
    starts = [i for i, o in enumerate(out) if o[0] == first]
    seg = out[starts[-2]:starts[-1]] if len(starts) >= 2 else out
    total = sum(o[3] for o in seg)
    print("| kernel | grid | block | time (us) | share |\n|---|---|---|---|---|")
    for name, g, b, t in seg:
        short = name.replace("qsae::<unnamed>::", "").replace("void ", "").split("(")[0]
        print(f"| `{short}` | {g} | {b} | {t:.1f} | {100 * t / total:.1f} % |")
    print(f"| **sum** | | | **{total:.1f}** | |")


def ncu(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    for v in rows[2:]:
        print(f"kernel: `{v[h.index('Kernel Name')][:120]}`\n\n| metric | value |\n|---|---|")
        for m in METRICS:
            if m in h:
                print(f"| {m} | {v[h.index(m)]} {units[h.index(m)]} |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3:])
    else:
        ncu(sys.argv[2])
