"""per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv"""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    name = re.sub(r"\(.*", "", r[ik])
    tot[name][0] += 1; tot[name][1] += us
allus = sum(v[1] for v in tot.values())
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {n:5d} launches {us:10.1f} us total {us / n:8.2f} us each {100 * us / allus:5.1f} %")
print(f"{'all':40s} {sum(v[0] for v in tot.values()):5d} launches {allus:10.1f} us")
