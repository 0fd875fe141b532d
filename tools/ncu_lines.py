"""Per CUDA source line: stall samples and executed instructions from `ncu --page source --csv --print-source cuda,sass`.
   python tools/ncu_lines.py <csv> [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = defaultdict(lambda: [0, 0, ""])
hdr = None
tot_s = tot_i = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        # two "Source" columns: first = CUDA line text, second = SASS
        src_idx = [i for i, n in enumerate(r) if n == "Source"]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    try:
        smp = int(r[hdr["# Samples"]] or 0)
        ins = int(r[hdr["Instructions Executed"]] or 0)
    except ValueError:
        continue
    if r[hdr["Address"]] != "-":
        continue        # SASS rows repeat the per-line totals; keep the line rows only
    key = (cur_file, line)
    agg[key][0] += smp
    agg[key][1] += ins
    agg[key][2] = r[src_idx[0]].strip()[:110]
    tot_s += smp
    tot_i += ins
print("total samples %d, instructions %d" % (tot_s, tot_i))
byfile = defaultdict(lambda: [0, 0])
for (f, l), (s, i, t) in agg.items():
    byfile[f][0] += s
    byfile[f][1] += i
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s %6.2f %% samples  %6.2f %% instr" % (f, 100.0 * s / max(1, tot_s), 100.0 * i / max(1, tot_i)))
for (f, l), (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% s %5.2f%% i  %s:%d  %s" % (100.0 * s / max(1, tot_s), 100.0 * i / max(1, tot_i), f, l, t))
