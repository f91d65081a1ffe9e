#!/usr/bin/env python
"""Top stall lines of one kernel from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu_source_top.py file.csv [section_index] [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
print("sections:", [(k, rows[i][1][:60]) for k, i in enumerate(starts)])
a = starts[sec]; b = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
hdr = rows[a + 1]; body = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
iS, iA, iI = col["Source"], col["Warp Stall Sampling (All Samples)"], col["Instructions Executed"]
iW, iWI = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"]
num = lambda x: int(float(x)) if x not in ("", "-") else 0
tot = sum(num(r[iA]) for r in body); ti = sum(num(r[iI]) for r in body)
print("kernel:", rows[a][1][:100]); print("samples", tot, "warp instructions", ti)
order = sorted(range(len(body)), key=lambda k: -num(body[k][iA]))
for k in order[:top]:
    r = body[k]
    print(f"{k:5d} {num(r[iA]):6d} {100*num(r[iA])/max(tot,1):5.1f}% inst={num(r[iI]):>9d} shw={num(r[iW]):>8d}/{num(r[iWI]):>8d}  {r[iS][:120]}")
