"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import collections
import csv
import sys


def summarise(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        k = row["Kernel Name"]
        grid = row.get("Grid Size", "")
        a = agg.setdefault(k, [0, 0.0, grid])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["%-44s %6s %12s %8s  %s" % ("kernel", "n", "total_us", "share", "grid(first)")]
    for k, (n, t, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-44s %6d %12.1f %7.1f%%  %s" % (k[:44], n, t, 100 * t / tot, g))
    out.append("%-44s %6s %12.1f" % ("TOTAL", "", tot))
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1]))
