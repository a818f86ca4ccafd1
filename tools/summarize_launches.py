"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total / average
duration and share.  usage: summarize_launches.py launches.csv [skip_name_substring ...]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = sys.argv[2:]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = {h: i for i, h in enumerate(rows[hdr])}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= H["Metric Value"] or r[H["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[H["Kernel Name"]]
    if any(s in name for s in skip):
        continue
    unit = r[H["Metric Unit"]]
    v = float(r[H["Metric Value"]].replace(",", ""))
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    short = name.split("(")[0].replace("void ", "").replace("es::<unnamed>::", "")[:60]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values()) or 1.0
print("kernel | launches | total us | avg us | share   (ncu: cold cache, serialised -- compare SHARES)")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} {n:5d} {us:9.1f} {us / n:8.1f} {us / tot:6.3f}")
