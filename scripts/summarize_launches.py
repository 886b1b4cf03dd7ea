"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of the captured region.  Usage: python scripts/summarize_launches.py launches.csv [out.md]"""
import csv, io, re, sys
from collections import defaultdict

path = sys.argv[1]
lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    name = re.sub(r"^.*::", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v * scale; cnt[name] += 1
total = sum(tot.values())
out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for k in sorted(tot, key=tot.get, reverse=True):
    out.append(f"| {k} | {cnt[k]} | {tot[k]:.1f} | {100 * tot[k] / total:.1f}% |")
out.append(f"| **all** | {sum(cnt.values())} | {total:.1f} | 100% |")
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
