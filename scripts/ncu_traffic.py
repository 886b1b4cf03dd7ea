"""dram bytes per launch of the hot operators from an `ncu --set full` report -> profiles/r02_traffic.json
usage: python scripts/ncu_traffic.py rep.ncu-rep out.json   (kernels in the order scripts/prof_ops.py launches them)"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    i = idx[name]
    return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)


acc = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    key = ("ffn" if "k_ffn_tc" in name else "attn" if "k_attn_tc" in name else
           "outproj_ln" if "k_gemm_tc<256, 2, 1" in name else "qkv" if "k_gemm_tc<256, 2, 0" in name else None)
    if key is None:
        continue
    acc.setdefault(key, []).append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
res = {"source": "ncu --set full --clock-control none, B=256, S_ctx=77, one operator per launch (scripts/prof_ops.py): "
                 "dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches",
       "unit": "bytes per launch"}
for k, v in acc.items():
    res[k] = int(sum(v) / len(v))
json.dump(res, open(out, "w"), indent=1)
print(res)
