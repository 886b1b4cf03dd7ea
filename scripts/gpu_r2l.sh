#!/bin/bash
mkdir -p gpurun_out
for nob in 2 4; do for sn in 0 1; do
echo "== NOB=$nob SNAKE=$sn: attention shapes"
MLDB_ATTN_NOB=$nob MLDB_SNAKE=$sn timeout 600 python scripts/dbg_attn.py > gpurun_out/dbg_attn_l_${nob}_${sn}.log 2>&1; grep -c "finite=True" gpurun_out/dbg_attn_l_${nob}_${sn}.log; grep -v "finite=True" gpurun_out/dbg_attn_l_${nob}_${sn}.log | head -12
echo "== NOB=$nob SNAKE=$sn: whole-path tests"
MLDB_ATTN_NOB=$nob MLDB_SNAKE=$sn timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "whole_path or scheduling or product_path" > gpurun_out/pytest_l_${nob}_${sn}.log 2>&1; tail -3 gpurun_out/pytest_l_${nob}_${sn}.log | cut -c1-200
done; done
