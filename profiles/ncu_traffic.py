"""DRAM traffic of one kernel launch from an `ncu --set full` capture -> JSON that bench.py reads for roofline.traffic.
usage: python profiles/ncu_traffic.py report.ncu-rep kernel_regex out.json"""
import csv
import json
import subprocess
import sys

rep, rx, dst = sys.argv[1:4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
import re
row = [r for r in rows[2:] if re.search(rx, r[ki])][0]


def val(key):
    i = h.index(key)
    v = float(row[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


rec = {"kernel": row[ki].split("(")[0].strip(), "report": rep.split("/")[-1],
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "gpu_time_ms_under_ncu": float(row[h.index("gpu__time_duration.sum")].replace(",", "")) *
       {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[h.index("gpu__time_duration.sum")].lower(), 1.0)}
rec["dram_bytes"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
json.dump(rec, open(dst, "w"), indent=1)
print(json.dumps(rec))
