"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv --log-file X.csv` launch list (or
the raw page of an .ncu-rep) into a per-kernel markdown table for profiles/: launches, total / mean device time, DRAM bytes.

    python scripts/launch_summary.py gpurun_out/launches.csv profiles/r02_launches_grid_b1.md "title line"
"""
import csv, io, sys
from collections import OrderedDict

src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = [l for l in open(src, errors="replace").read().splitlines() if l.startswith('"')]
rows = list(csv.reader(io.StringIO("\n".join(lines))))
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ii = hdr.index("ID")
SC = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = OrderedDict()
seen = {}
for r in rows[1:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("vlfm::", "")
    d = per.setdefault(name, {"n": 0, "us": 0.0, "bytes": 0.0})
    v = float(r[vi].replace(",", "")) * SC.get(r[ui], 1.0)
    if r[mi] == "gpu__time_duration.sum":
        d["us"] += v
        if (r[ii], name) not in seen:
            seen[(r[ii], name)] = 1
            d["n"] += 1
    elif r[mi].startswith("dram__bytes"):
        d["bytes"] += v
tot = sum(d["us"] for d in per.values())
n = sum(d["n"] for d in per.values())
with open(out, "w") as f:
    f.write(f"# {title}\n\n{n} launches, {tot:.1f} us total device time (ncu: cold cache, serialised -- compare shares, not absolutes)\n\n")
    f.write("| kernel | launches | total us | share | mean us | DRAM MB |\n|---|---|---|---|---|---|\n")
    for name, d in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"| {name[:70]} | {d['n']} | {d['us']:.1f} | {100 * d['us'] / max(tot, 1e-9):.1f}% | {d['us'] / max(d['n'], 1):.2f} | {d['bytes'] / 1e6:.2f} |\n")
print(open(out).read())
