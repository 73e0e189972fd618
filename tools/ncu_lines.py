#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an ncu report captured with --import-source on.
usage: ncu_lines.py report.ncu-rep <kernel substring> [top N]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = defaultdict(lambda: [0, 0, 0, ""])
cur_file = cur_fn = None
hdr = None
seen_fn = set()
use = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        use = pat in cur_fn
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for j, h in enumerate(hdr):
            ix.setdefault(h, j)
        continue
    if not use or hdr is None or len(r) != len(hdr):
        continue
    try:
        key = (cur_file, int(r[0]))
        a = agg[key]
        a[0] += int(r[ix["Instructions Executed"]])
        a[1] += int(r[ix["# Samples"]])
        a[2] += int(r[ix["Thread Instructions Executed"]])
        a[3] = r[1].strip()[:120]
    except ValueError:
        pass
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"kernel ~ {pat}: warp instr {ti}, samples {ts} (a kernel captured k times is counted k times)")
for key, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{100*a[0]/ti:5.1f}% inst {100*a[1]/ts:5.1f}% stall-samples  thr/inst {a[2]/max(1,a[0]):4.1f}  {key[0]}:{key[1]}  {a[3]}")
