"""Top stall locations of one kernel from an .ncu-rep (source page, SASS).
usage: python profiles/ncu_hot.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ci, si, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[2:]:
    if len(r) <= si or r[0] in ("Kernel Name", "Address"):
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    st = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
    data.append((n, int(r[ie] or 0), r[ci].strip()[:70], st))
tot = sum(d[0] for d in data)
print("total samples", tot, " SASS lines", len(data), " warp-instr", sum(d[1] for d in data))
for n, ex, s, st in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% exec=%9d  %-70s %s" % (n, 100.0 * n / max(tot, 1), ex, s, st))
