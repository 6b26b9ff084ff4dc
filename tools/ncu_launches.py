"""Summarise an `ncu --csv` launch list: per (kernel, grid) count, total time, share, and optional metrics."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        key = (r["ID"], r["Kernel Name"][:48], r["Grid Size"], r["Block Size"])
        per.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "") or 0)
    agg = collections.OrderedDict()
    for (_, name, grid, block), m in per.items():
        a = agg.setdefault((name, grid, block), collections.defaultdict(float))
        a["n"] += 1
        for k, v in m.items():
            a[k] += v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print("%-50s %-16s %4s %10s %6s %8s  other metrics (mean)" % ("kernel", "grid", "n", "total us", "share", "avg us"))
    for (name, grid, block), a in agg.items():
        t = a["gpu__time_duration.sum"]
        extra = " ".join("%s=%.1f" % (k.split(".")[0].replace("sm__", "").replace("smsp__", "").replace("launch__", ""), v / a["n"])
                         for k, v in a.items() if k not in ("n", "gpu__time_duration.sum"))
        print("%-50s %-16s %4d %10.1f %5.1f%% %8.1f  %s" % (name, grid, a["n"], t / 1e3, 100 * t / tot, t / a["n"] / 1e3, extra))
    print("total %.3f ms" % (tot / 1e6))


if __name__ == "__main__":
    main(sys.argv[1])
