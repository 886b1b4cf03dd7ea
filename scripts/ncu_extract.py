"""Extract the roofline-relevant metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page
raw --csv`) into a small markdown table.  Usage: python scripts/ncu_extract.py rep.ncu-rep out.md"""
import csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
cols = {
    "Kernel Name": "kernel", "Grid Size": "grid", "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram rd", "dram__bytes_write.sum": "dram wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe % (active)",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor pipe % (elapsed)",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram %",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "L2 %",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy %",
    "launch__registers_per_thread": "regs",
}
idx = {h: i for i, h in enumerate(hdr)}
units = rows[1]
lines = ["| " + " | ".join(cols.values()) + " |", "|" + "---|" * len(cols)]
for r in rows[2:]:
    vals = []
    for k in cols:
        i = idx.get(k)
        v = r[i] if i is not None else ""
        if k == "Kernel Name":
            v = v.split("(")[0].replace("void <unnamed>::", "")
        elif i is not None and units[i]:
            v = f"{v} {units[i]}"
        vals.append(v)
    lines.append("| " + " | ".join(vals) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
