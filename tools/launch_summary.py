"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / mean / total and share."""
import collections
import csv
import sys


def main(path, skip=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    rows = list(csv.DictReader(lines))[skip:]
    for x in rows:
        k = x["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:48]
        agg.setdefault(k, []).append(float(x["Metric Value"].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':48s} {'n':>5s} {'mean_us':>10s} {'min_us':>9s} {'max_us':>9s} {'total_us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:48s} {len(v):5d} {sum(v)/len(v)/1e3:10.2f} {min(v)/1e3:9.2f} {max(v)/1e3:9.2f} {sum(v)/1e3:10.1f} {100*sum(v)/tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
