"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
        k = row["Kernel Name"][:90]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    print("# %s : %d launches, %.2f ms total device time (cold-cache, serialised: compare shares)" % (path, sum(a[0] for a in agg.values()), tot / 1e6))
    print("%-92s %7s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if t / tot < 0.0005:
            continue
        print("%-92s %7d %12.2f %10.1f %6.1f%%" % (k, n, t / 1e6, t / n / 1e3, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
