"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        if unit in ("nsecond", "ns"):
            v /= 1e3
        elif unit in ("msecond", "ms"):
            v *= 1e3
        elif unit in ("second", "s"):
            v *= 1e6
        tot[name][0] += 1
        tot[name][1] += v
total = sum(t for _, t in tot.values())
print(f"{'kernel':44s} {'launches':>9s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
    print(f"{k:44s} {c:9d} {t / 1e3:10.2f} {t / c:9.1f} {100 * t / total:6.1f}%")
print(f"{'TOTAL':44s} {sum(c for c, _ in tot.values()):9d} {total / 1e3:10.2f}")
