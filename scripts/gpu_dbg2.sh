#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dbg.log
for d in 0 8 16 1; do
  echo "== MLDB_TC_DBG=$d" >> gpurun_out/dbg.log
  MLDB_TC_DBG=$d timeout 120 python scripts/prof_ops.py qkv ffn1 2>&1 | tail -2 >> gpurun_out/dbg.log
done
cat gpurun_out/dbg.log
