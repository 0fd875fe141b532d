"""Summarise gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and gpurun_out/prof.ncu-rep (ncu --set full)
into small tracked text files under profiles/.   usage: python tools/summarize_ncu.py <tag>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []
lp = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.isfile(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    cols, data = rows[h], rows[h + 1:]
    ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e6 if r[ui] == "ns" else (v / 1e3 if r[ui] == "us" else v)   # -> ms
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out.append("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)")
    out.append("# command: ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline")
    out.append("launches=%d total_ms=%.3f" % (len(data), tot))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-72s n=%5d %12.3f ms %6.2f%%" % (n, a[0], a[1], 100 * a[1] / tot))
rp = os.path.join(ROOT, "gpurun_out", sys.argv[2] if len(sys.argv) > 2 else "prof.ncu-rep")
if os.path.isfile(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__inst_executed.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
            "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
    out.append("")
    out.append("# ncu --set full --clock-control none (one capture per kernel; dram bytes per launch)")
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        out.append("kernel: " + r[ki][:110])
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                out.append("    %-75s %16s %s" % (w, r[i], units[i]))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
path = os.path.join(ROOT, "profiles", "%s_ncu_summary.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print(path, len(out), "lines")
