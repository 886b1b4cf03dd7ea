#!/bin/bash
for bn in 256 128; do for d in 0 2; do
  echo "== BN=$bn DBG=$d"; MLDB_TC_BN=$bn MLDB_TC_DBG=$d timeout 120 python scripts/prof_ops.py qkv ffn1 2>&1 | tail -2
done; done
