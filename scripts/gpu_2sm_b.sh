#!/bin/bash
mkdir -p gpurun_out
for cfg in "MLDB_TC_2SM=0" "MLDB_TC_2SM=1" "MLDB_TC_2SM=1 MLDB_TC_DBG=8"; do
  echo "== $cfg"
  env $cfg timeout 100 python scripts/prof_ops.py qkv outproj_ln ffn1 2>&1 | tail -3
done
MLDB_TC_2SM=1 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_2smb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])"
