"""Aggregate the output of ncu_lines.py (all lines) by kernel phase.
usage: python profiles/ncu_lines.py rep rx obj mangled 100000 outer | python profiles/ncu_phases.py file.inc name:first:last ..."""
import re
import sys

fname = sys.argv[1]
ph = [(a.split(":")[0], int(a.split(":")[1]), int(a.split(":")[2])) for a in sys.argv[2:]]
agg = {}
for l in sys.stdin:
    m = re.match(r"\s*([\d.]+)% smp\s+([\d.]+)% ins\s+(\S+):(\d+)", l)
    if not m:
        continue
    s, i, f, ln = float(m[1]), float(m[2]), m[3], int(m[4])
    key = "other:" + f
    if f == fname:
        for n, a, b in ph:
            if a <= ln <= b:
                key = n
    a = agg.setdefault(key, [0, 0])
    a[0] += s
    a[1] += i
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-28s smp %5.1f%%  ins %5.1f%%" % (k, v[0], v[1]))
