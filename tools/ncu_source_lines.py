#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).

    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep render_backward_kernel [top N]

Prints, for the N source lines with the most warp-stall samples, their share of executed warp
instructions, of stall samples, average active threads and the dominant stall reasons.
"""
import csv
import subprocess
import sys


def main(rep, kernel, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    cur_file = ""
    recs = []
    seen_kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if seen_kernel is None:
                seen_kernel = r[1]
            elif r[1] != seen_kernel:
                pass  # several launches of the same kernel are listed one after another
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        g = lambda name: r[hdr.index(name)]
        try:
            inst = int(g("Instructions Executed"))
            samp = int(g("# Samples"))
            thr = int(g("Thread Instructions Executed"))
        except ValueError:
            continue
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "(Not Issued)" not in h:
                try:
                    v = int(r[i])
                except (ValueError, IndexError):
                    v = 0
                if v:
                    stalls[h[6:]] = v
        recs.append((cur_file, ln, r[1].strip(), inst, samp, thr, stalls))
    # merge duplicate (file, line) entries (same line inlined at several places / several launches)
    merged = {}
    for f, ln, src, inst, samp, thr, st in recs:
        m = merged.setdefault((f, ln), [src, 0, 0, 0, {}])
        m[1] += inst
        m[2] += samp
        m[3] += thr
        for k, v in st.items():
            m[4][k] = m[4].get(k, 0) + v
    tot_i = sum(m[1] for m in merged.values()) or 1
    tot_s = sum(m[2] for m in merged.values()) or 1
    print(f"kernel {kernel}: {tot_i} warp instructions, {tot_s} stall samples (all captured launches)")
    print(f"{'file:line':28s} {'inst%':>6s} {'samp%':>6s} {'thr/warp':>8s}  top stalls | source")
    for (f, ln), m in sorted(merged.items(), key=lambda kv: -kv[1][2])[:top]:
        st = sorted(m[4].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k}:{100 * v / max(m[2], 1):.0f}%" for k, v in st)
        avg = m[3] / m[1] if m[1] else 0
        print(f"{f + ':' + str(ln):28s} {100 * m[1] / tot_i:6.2f} {100 * m[2] / tot_s:6.2f} {avg:8.1f}  {sts} | {m[0][:80]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
