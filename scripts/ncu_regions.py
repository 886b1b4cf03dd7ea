"""Summarise an ncu --set full report: per kernel, stall samples per source line (top lines)."""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; out = []
for r in rows:
    if r and r[0] in ("Address", "#"):
        if hdr is not None: break
        hdr = r; continue
    if hdr and len(r) >= len(hdr) - 2: out.append(r)
print(hdr[:8])
i_s = hdr.index("Warp Stall Sampling (All Samples)"); i_src = hdr.index("Source")
tot = sum(int(r[i_s] or 0) for r in out)
print("total samples", tot, "rows", len(out))
for i, r in sorted(sorted(enumerate(out), key=lambda t: -int(t[1][i_s] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]):
    print(i, r[i_s], r[i_src].strip()[:110])
