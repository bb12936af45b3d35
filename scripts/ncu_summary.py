"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small CSV + markdown table
for profiles/: duration, DRAM bytes, DRAM %, tensor-pipe %, L2 %, registers, grid, achieved bandwidth."""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
# input: an .ncu-rep, or the CSV of its raw page (`ncu -i X.ncu-rep --page raw --csv > X.csv`, exported on the GPU box: reports are
# too large to bring back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
raw = "\n".join(l for l in raw.splitlines() if l.startswith('"'))
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dsmem")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
recs = []
for r in rows[2:]:
    d = {}
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            d[n] = v.split("(")[0].replace("void ", "").replace("vlfm::", "")
        else:
            try:
                d[n] = float(v.replace(",", "")) * SC.get(units[i], 1)
            except ValueError:
                d[n] = v
    d["dram_gbs"] = (d.get("dram_rd", 0) + d.get("dram_wr", 0)) / max(d.get("dur", 1), 1e-9) / 1e3  # bytes/us -> GB/s
    recs.append(d)
cols = ["kernel", "dur", "grid", "block", "regs", "dsmem", "dram_rd", "dram_wr", "dram_gbs", "dram_pct", "l2_pct", "tensor_pct", "sm_pct", "occ_pct"]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f); w.writerow(cols)
    for d in recs: w.writerow([d.get(c, "") for c in cols])
with open(out + ".md", "w") as f:
    f.write("| kernel | dur us | grid x block | regs | dyn smem | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % | L2 % | tensor % | occ % |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for d in recs:
        f.write(f"| {d['kernel'][:60]} | {d.get('dur',0):.2f} | {int(d.get('grid',0))}x{int(d.get('block',0))} | {int(d.get('regs',0))} | {int(d.get('dsmem',0))} | "
                f"{d.get('dram_rd',0)/1e6:.2f} | {d.get('dram_wr',0)/1e6:.2f} | {d['dram_gbs']:.0f} | {d.get('dram_pct',0):.1f} | {d.get('l2_pct',0):.1f} | {d.get('tensor_pct',0):.1f} | {d.get('occ_pct',0):.1f} |\n")
print(open(out + ".md").read())
