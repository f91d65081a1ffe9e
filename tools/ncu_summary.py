#!/usr/bin/env python
"""Markdown table of the key metrics of an `ncu --set full` report (one column per kernel, first launch of each)
and, with --traffic CONFIG, the DRAM bytes per launch merged into profiles/ncu_traffic.json.

    ncu -i gpurun_out/prof_X.ncu-rep --page raw --csv > /tmp/prof_X.csv
    python tools/ncu_summary.py /tmp/prof_X.csv [--traffic c3 --source profiles/r01_ncu_step_c3.md --R 14200000]
"""
import argparse
import csv
import json
from pathlib import Path

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__grid_size",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
KERNELS = ["preprocess_kernel", "emit_instances_kernel", "onesweep_pass_kernel", "radix_hist_kernel", "render_forward",
           "render_backward", "gaussian_backward_kernel", "pixel_loss_forward_kernel", "match_kernel<(int)0>",
           "match_kernel<(int)1>", "knn_query_kernel"]
TRAFFIC_KEYS = {"render_backward": "render_backward_kernel", "render_forward": "render_forward_kernel",
                "preprocess_kernel": "preprocess_kernel", "gaussian_backward_kernel": "gaussian_backward_adam_kernel"}


def scale(v, unit):
    u = unit.lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic")
    ap.add_argument("--source", default="")
    ap.add_argument("--R", type=int, default=0)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    cols = []
    for name in KERNELS:
        for r in data:
            if name in r[ki]:
                cols.append((name, r))
                break
    print("| metric | " + " | ".join(n for n, _ in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for w in WANT:
        if w not in hdr:
            continue
        i = hdr.index(w)
        print(f"| {w} [{units[i]}] | " + " | ".join(f"{float(r[i].replace(',', '')):.4g}" for _, r in cols) + " |")
    if a.traffic:
        tj = Path(__file__).resolve().parent.parent / "profiles" / "ncu_traffic.json"
        tab = json.loads(tj.read_text())
        cfg = tab.setdefault(a.traffic, {})
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for n, r in cols:
            if n in TRAFFIC_KEYS:
                cfg[TRAFFIC_KEYS[n]] = {"dram_read_bytes": int(scale(float(r[ir]), units[ir])),
                                        "dram_write_bytes": int(scale(float(r[iw]), units[iw])),
                                        "source": a.source, "R": a.R}
        tj.write_text(json.dumps(tab, indent=1) + "\n")


if __name__ == "__main__":
    main()
