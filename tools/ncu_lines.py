#!/usr/bin/env python
"""Per-source-line instruction and stall totals of one kernel.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv; ncu_lines.py f.csv <kernel substring> [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = []   # (file, line, text, inst, samples)
i = 0
cur_file = None; cur_fn = None; hdr = None
seen_fn = set()
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": cur_file = r[1]
    elif r and r[0] == "Function Name": cur_fn = r[1]
    elif r and r[0] == "Line No": hdr = {h: k for k, h in enumerate(r)}; hdr_list = r
    elif hdr and cur_fn and want in cur_fn and len(r) == len(hdr_list) and r[0] not in ("", "-"):
        # first 'Source' column is index 1; instructions executed etc.
        try:
            inst = int(float(r[hdr["Instructions Executed"]] or 0)); smp = int(float(r[hdr["Warp Stall Sampling (All Samples)"]] or 0))
        except ValueError:
            inst = smp = 0
        out.append((cur_file.split("/")[-1], int(r[0]), r[1].strip()[:100], inst, smp, cur_fn[:40]))
    i += 1
# several launches of the same kernel may be present: keep the first launch per (file, line)
agg = {}
for f, ln, txt, inst, smp, fn in out:
    k = (f, ln)
    if k not in agg: agg[k] = [txt, inst, smp]
ti = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"kernel ~{want}: {ti} warp instructions, {ts} samples (first launch per line)")
for (f, ln), (txt, inst, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:<5d} inst {100*inst/max(ti,1):5.1f}%  stall {100*smp/max(ts,1):5.1f}%  {txt}")
