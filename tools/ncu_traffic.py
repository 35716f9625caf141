"""profiles/traffic.json from one `ncu --set full` capture of the soil kernel: DRAM bytes and executed FP64 flops per site-day, tagged with the hash
of the library build the capture was taken on (bench.py refuses a capture of another build).

usage: ncu_traffic.py raw.csv sass.csv SITE_DAYS LIBRARY_HASH [out.json]
  raw.csv   `ncu -i rep --page raw --csv`      (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the captured launch)
  sass.csv  `ncu -i rep --page source --csv`   (per-instruction "Thread Instructions Executed": DADD + DMUL + 2 DFMA = executed FP64 flops)
"""
import csv
import json
import os
import sys

raw, sass, site_days, lib_hash = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def to_bytes(key):
    v, u = m[key]
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
import re
srows = list(csv.reader(open(sass)))
hi = next(i for i, r in enumerate(srows) if "Source" in r and "Thread Instructions Executed" in r)
ix = {h: i for i, h in enumerate(srows[hi])}
src_col, tcol = ix["Source"], ix["Thread Instructions Executed"]
ops = {"DFMA": 0, "DADD": 0, "DMUL": 0}
total = 0
for r in srows[hi + 1:]:
    if len(r) <= tcol or not r[tcol].isdigit():
        continue
    n = int(r[tcol])
    total += n
    mm = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[src_col])      # opcode after an optional predicate
    if mm and mm.group(1) in ops:
        ops[mm.group(1)] += n
flops = ops["DADD"] + ops["DMUL"] + 2 * ops["DFMA"]
res = {"library_hash": lib_hash, "site_days_in_capture": site_days, "kernel": m.get("Kernel Name", ("?",))[0],
       "gpu_time_ms": float(m["gpu__time_duration.sum"][0].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(m["gpu__time_duration.sum"][1], 1e-6),
       "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
       "dram_bytes_per_site_day": dram / site_days, "thread_inst_per_site_day": total / site_days,
       "fp64_thread_inst_per_site_day": {k: v / site_days for k, v in ops.items()}, "fp64_flop_per_site_day": flops / site_days}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
