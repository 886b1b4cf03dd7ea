"""SASS evidence per kernel of libmldb200.so: tcgen05 (UTC*MMA), TMEM (LDTM/STTM), TMA (UTMALDG/UTMASTG),
legacy tensor path (HMMA), MUFU.  usage: python scripts/sass_summary.py [lib.so] > profiles/rNN_sass_summary.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "mld_b200/libmldb200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG",
       "UTMASTG": r"\bUTMASTG", "UTCBAR": r"\bUTCBAR", "SYNCS": r"\bSYNCS", "HMMA": r"\bHMMA", "MUFU": r"\bMUFU", "LDGSTS": r"\bLDGSTS"}
kern, counts, ninstr = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\((CUtensorMap_st|ActBuf|float|int|long|AttnArgs|GemmArgs|LnArgs|__half).*", "", kern)
        while kern in counts: kern += "'"          # static kernels compiled into more than one translation unit
        counts[kern] = collections.Counter(); ninstr[kern] = 0
        continue
    if kern and re.search(r"/\*[0-9a-f]{4,}\*/", line):
        ninstr[kern] += 1
        for k, p in pat.items():
            if re.search(p, line): counts[kern][k] += 1
cols = list(pat)
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for k, c in counts.items():
    print(f"| `{k}` | {ninstr[k]} | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
tot = collections.Counter()
for c in counts.values(): tot.update(c)
print("| **all** | %d | " % sum(ninstr.values()) + " | ".join(str(tot[x]) for x in cols) + " |")
