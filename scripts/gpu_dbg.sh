#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 6; do
  echo "== MLDB_TC_DBG=$d" >> gpurun_out/dbg.log
  MLDB_TC_DBG=$d timeout 120 python scripts/prof_ops.py qkv ffn1 ffn2_ln outproj_ln 2>&1 | tail -4 >> gpurun_out/dbg.log
done
cat gpurun_out/dbg.log
