"""Summarise the source page of an ncu report (--import-source on): stall samples per SASS opcode class and the hottest
instructions.   python tools/ncu_src.py <source.csv> [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        smp = int(r[ix["# Samples"]] or 0)
        ins = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    data.append((r[ix["Source"]], smp, ins, {s: int(r[ix[s]] or 0) for s in stalls}, r))
tot_s = sum(d[1] for d in data)
tot_i = sum(d[2] for d in data)
print("total samples %d, warp instructions executed %d" % (tot_s, tot_i))
agg = defaultdict(lambda: [0, 0])
st_tot = defaultdict(int)
for src, smp, ins, st, r in data:
    op = src.split()[0] if src.split() else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0] + ("." + src.split()[0].split(".")[1] if "." in src.split()[0] and op in ("MUFU", "LDS", "STS", "HADD2", "F2FP", "LDG", "STG") else "")
    agg[op][0] += smp
    agg[op][1] += ins
    for k, v in st.items():
        st_tot[k] += v
print("-- stall reasons (all samples)")
for k, v in sorted(st_tot.items(), key=lambda kv: -kv[1])[:12]:
    print("   %-26s %6.2f %%" % (k, 100.0 * v / max(1, tot_s)))
print("-- by opcode: %samples  %instructions")
for op, (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("   %-14s %6.2f %%  %6.2f %%" % (op, 100.0 * smp / max(1, tot_s), 100.0 * ins / max(1, tot_i)))
print("-- hottest instructions")
for src, smp, ins, st, r in sorted(data, key=lambda d: -d[1])[:top]:
    why = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("   %5.2f %%  %-70s %s" % (100.0 * smp / max(1, tot_s), src[:70], " ".join("%s=%d" % (k[6:], v) for k, v in why)))
