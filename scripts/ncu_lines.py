"""Stall samples per CUDA source line of one kernel of an ncu --set full --import-source on report.
usage: python scripts/ncu_lines.py rep.ncu-rep ::kernel_name:invocation [top]"""
import csv, io, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the report is a sequence of per-file sections: a "File Path" row, a header row, then one row per CUDA line
tot = collections.Counter(); reasons = collections.defaultdict(collections.Counter); src = {}
hdr = None; fpath = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    try: line = int(r[0])
    except ValueError: continue
    i_s = hdr.index("Warp Stall Sampling (All Samples)")
    try: s = int(r[i_s] or 0)
    except ValueError: continue
    if not s: continue
    key = (fpath.split("/")[-1], line)
    tot[key] += s; src[key] = r[1].strip()[:110]
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0"): reasons[key][h] += int(r[i])
allS = sum(tot.values())
print("total samples", allS)
for key, s in tot.most_common(top):
    rs = ", ".join(f"{h[6:]}={v}" for h, v in reasons[key].most_common(2))
    print(f"{key[0]}:{key[1]:5d} {s:6d} {100*s/allS:5.1f}%  {src[key]}   [{rs}]")
