#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel device time of
the LAST complete step (a step ends with the last fused-Adam launch or, with Adam applied inside the
backward, with gaussian_backward_kernel), as shares of the step.

    python tools/launch_summary.py gpurun_out/launches_r01.csv [> profiles/r01_launches.md]
"""
import collections
import csv
import re
import sys


def main(path, adam_per_step=6):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    adam = [i for i, x in enumerate(rows) if "adam" in x["Kernel Name"]]
    ends = [adam[i] for i in range(adam_per_step - 1, len(adam), adam_per_step)]
    if not adam:  # optimizer-in-backward: the per-Gaussian backward kernel is the step's last launch
        ends = [i for i, x in enumerate(rows) if "gaussian_backward" in x["Kernel Name"]]
    if len(ends) < 2:
        raise SystemExit("need at least two complete steps in the launch list")
    s, e = ends[-2] + 1, ends[-1] + 1
    agg = collections.OrderedDict()
    tot = 0.0
    for x in rows[s:e]:
        n = re.sub(r"\(.*", "", x["Kernel Name"])
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"<.*", "", n) if n.startswith("w3d::") else n[:100]
        t = float(x["Metric Value"]) / 1e3
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    ours = sum(t for n, (c, t) in agg.items() if n.startswith("w3d::"))
    print(f"# launch list summary of `{path}`")
    print(f"\nlast complete step: {e - s} launches, {tot:.1f} us of serialised cold-cache device time "
          f"({len(rows)} launches captured); our kernels (w3d::*) {ours:.1f} us = {100 * ours / tot:.1f}%\n")
    print("| device us | launches | share | kernel |")
    print("|---:|---:|---:|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {t:.1f} | {c} | {100 * t / tot:.1f}% | `{n}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 6)
