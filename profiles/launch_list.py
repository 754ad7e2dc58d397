"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file x.csv): per kernel launches, total
and mean duration, share of the profiled GPU time.   usage: python profiles/launch_list.py x.csv [own_only]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
own = len(sys.argv) > 2
agg = collections.OrderedDict()
for r in rows:
    name = r[4]
    if own and ("at::" in name or "nccl" in name.lower()):
        continue
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e6
tot = sum(a[1] for a in agg.values())
print("%-72s %7s %10s %10s %6s" % ("kernel", "count", "total ms", "mean ms", "share"))
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %7d %10.3f %10.4f %5.1f%%" % (k, n, ms, ms / n, 100 * ms / tot))
print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
