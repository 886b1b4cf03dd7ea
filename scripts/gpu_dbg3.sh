#!/bin/bash
for fp in 1 0; do for d in 0 1 2 32 34 38; do
  echo "== FFN_PAIR=$fp DBG=$d"; MLDB_FFN_PAIR=$fp MLDB_TC_DBG=$d timeout 120 python scripts/prof_ops.py ffn 2>&1 | tail -1
done; done
