#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05.mma microbenchmark"
for op in attn outproj_ln qkv ffn; do echo "== timeline $op"; timeout 200 python scripts/timeline.py $op 120 > gpurun_out/timeline_$op.txt 2>&1; head -60 gpurun_out/timeline_$op.txt | cut -c1-300; done
