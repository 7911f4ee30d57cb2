#!/usr/bin/env python
"""Aggregate an `ncu -i rep --page source --print-source cuda,sass --csv` export by CUDA source line.

    python scripts/ncu_lines.py file.csv [top] [file-filter]

Every SASS row is attributed to the CUDA line it is listed under; per line: warp instructions
executed, stall samples, shared-memory wavefronts and the two dominant stall reasons."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur_file, hdr, cur = None, None, None
agg = collections.defaultdict(lambda: {"inst": 0, "samp": 0, "wf": 0, "src": "", "st": collections.Counter()})
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        n = len(r)
        stall_cols = [(h, i) for h, i in hdr.items() if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur = (cur_file, int(r[0]))
        agg[cur]["src"] = r[1].strip()[:80]
        continue
    if len(r) < 8 or cur is None:
        continue
    # SASS row; index from the end (source text with quotes can split into extra fields)
    def col(i):
        v = r[i - n]
        return int(v) if v.isdigit() else 0
    if not any(f.startswith("0x") for f in r[:6]):
        continue
    a = agg[cur]
    a["inst"] += col(hdr["Instructions Executed"])
    a["samp"] += col(hdr["# Samples"])
    a["wf"] += col(hdr["L1 Wavefronts Shared"])
    for h, i in stall_cols:
        a["st"][h[6:]] += col(i)
ti = sum(a["inst"] for a in agg.values()) or 1
ts = sum(a["samp"] for a in agg.values()) or 1
tw = sum(a["wf"] for a in agg.values()) or 1
tot_st = collections.Counter()
for a in agg.values():
    tot_st.update(a["st"])
print(f"total warp-inst {ti}  samples {ts}  smem wavefronts {tw}")
print("stalls:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in tot_st.most_common(9)))
flt = sys.argv[3] if len(sys.argv) > 3 else ""
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    if flt and flt not in f:
        continue
    st = " ".join(f"{k}:{100 * v / max(a['samp'], 1):.0f}" for k, v in a["st"].most_common(2))
    print(f"{f[:24]:24s}:{ln:4d} inst {100 * a['inst'] / ti:5.1f}%  samp {100 * a['samp'] / ts:5.1f}%  "
          f"wf {100 * a['wf'] / tw:5.1f}%  [{st}]  {a['src']}")
