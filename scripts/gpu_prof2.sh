#!/bin/bash
mkdir -p gpurun_out
PROF_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ffn_tc|k_gemm_tc" -c 4 -f -o gpurun_out/prof_ffn python scripts/prof_ops.py ffn qkv > gpurun_out/prof_ffn.log 2>&1
tail -3 gpurun_out/prof_ffn.log
ls -la gpurun_out/prof_ffn.ncu-rep
