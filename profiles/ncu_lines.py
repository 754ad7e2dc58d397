"""Attribute the SASS-level samples / executed instructions of one kernel in an .ncu-rep to CUDA source lines
(nvdisasm -g line tables of the matching cubin).  usage:
   python profiles/ncu_lines.py report.ncu-rep kernel_regex object.o mangled_substring [top_n]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, rx, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
by_ins = len(sys.argv) > 6 and "ins" in sys.argv[6]     # sort by executed instructions instead of samples
outer = len(sys.argv) > 6 and "outer" in sys.argv[6]    # attribute inlined code to the outermost call site (nvdisasm -gi)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
ins = []
for r in rows[2:]:
    if len(r) <= ie or not r[0].startswith("0x"):
        continue
    ins.append((int(r[0], 16), int(r[si] or 0), int(r[ie] or 0), [int(r[i] or 0) for i in stall_cols], r[1].strip()))
base = ins[0][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
# a linked library holds one cubin per translation unit: take the one whose symbol table names the kernel
cubs = sorted(glob.glob(os.path.join(tmp, "*.cubin")))
cub = next((c for c in cubs if mangled.encode() in open(c, "rb").read()), cubs[0])
dis = subprocess.run(["nvdisasm", "-gi" if outer else "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, infunc = None, False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        infunc = mangled in l
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        # with -gi a chain "X inlined at Y" precedes the instruction; its last line is the outermost frame
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
agg = {}
for addr, ns, ne, st, txt in ins:
    key = line_of.get(addr - base, ("?", 0))
    a = agg.setdefault(key, [0, 0, [0] * len(stall_cols)])
    a[0] += ns; a[1] += ne
    for i, v in enumerate(st):
        a[2][i] += v
tot_s, tot_e = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("samples", tot_s, "warp-instr", tot_e, "instructions", len(ins))
srcs = {}
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if by_ins else 0])[:top]:
    if f not in srcs:
        cand = (glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)) +
                glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "csrc", f)))
        srcs[f] = open(cand[0]).read().splitlines() if cand else []
    text = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    st = sorted(zip(a[2], [h[i][6:] for i in stall_cols]), reverse=True)[:2]
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %-80s %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_e, f, ln, text,
                                                        " ".join("%s=%d" % (n, v) for v, n in st)))
