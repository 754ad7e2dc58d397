"""Opcode mix of one kernel in an .ncu-rep, weighted by executed warp instructions and by stall samples.
usage: python profiles/ncu_opmix.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys
from collections import Counter

rep, rx = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
ex, sm = Counter(), Counter()
for r in rows[2:]:
    if len(r) <= iE:
        continue
    toks = r[iS].split()
    opc = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    opc = opc.rstrip(";")
    base = opc.split(".")[0]
    ex[base] += int(r[iE] or 0)
    sm[base] += int(r[iN] or 0)
te, ts = sum(ex.values()), sum(sm.values())
print("warp-instr", te, "samples", ts)
for k, v in ex.most_common(top):
    print("  %-10s %6.2f%% ins  %6.2f%% smp" % (k, 100.0 * v / te, 100.0 * sm[k] / max(ts, 1)))
