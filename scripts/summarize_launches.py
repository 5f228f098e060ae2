#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the driver-side evidence of
which kernels a command launches and what share of the time each takes).

    python scripts/summarize_launches.py gpurun_out/f_launches.csv > profiles/r02_launches_summary.txt
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    agg = collections.OrderedDict()
    n = 0
    for r in csv.reader(open(path)):
        if len(r) < 10 or not r[0].isdigit():
            continue
        n += 1
        name = re.sub(r"\(.*", "", r[4])[:84]
        v = float(r[-1].replace(",", ""))
        unit = r[-2]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "cfm::" in k)
    print(f"# {path}: {n} launches, {tot / 1e3:.2f} ms of kernel time; cfm:: kernels {100 * ours / tot:.1f} % of it "
          f"({sum(a[0] for k, a in agg.items() if 'cfm::' in k)} launches)")
    print(f"# {'kernel':<84} {'n':>5} {'total us':>11} {'share':>7} {'mean us':>9} {'min':>9} {'max':>9}")
    for k, (c, t, lo, hi) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<86} {c:>5} {t:>11.1f} {100 * t / tot:>6.1f}% {t / c:>9.1f} {lo:>9.1f} {hi:>9.1f}")


if __name__ == "__main__":
    main()
