"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step
(between two gather_relu launches), aggregated per kernel.  usage: launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def norm(n):
    n = n.replace("void ", "").replace("(anonymous namespace)::", "").replace("at::native::", "at::")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?", n)
    if not m:
        return n[:70]
    base, tmpl = m.group(1), (m.group(2) or "")
    if base.startswith("at::") or base.startswith("cub::"):
        tmpl = ""
    return (base + tmpl)[:70]


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines)]
    starts = [i for i, (n, _) in enumerate(rows) if "gather_relu" in n]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else min(2, len(starts) - 2)   # skip the allocating first steps
    step = rows[starts[which]:starts[which + 1]]
    tot = sum(t for _, t in step)
    agg = collections.OrderedDict()
    for n, t in step:
        k = norm(n)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += t
    print("launches/step %d, sum of kernel durations %.3f ms (ncu: cold-cache, serialised)" % (len(step), tot / 1e6))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-72s n=%3d %9.1f us %5.1f%%" % (k, c, t / 1000, 100 * t / tot))
    print("--- launches > 100 us, in order")
    for n, t in step:
        if t > 100e3:
            print("%-72s %9.1f us" % (norm(n), t / 1000))


if __name__ == "__main__":
    main(sys.argv[1])
