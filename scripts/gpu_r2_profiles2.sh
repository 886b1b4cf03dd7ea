#!/bin/bash
# second ncu pass: out-proj + LN and the fused FFN (the first pass's -c 8 ends after QKV and attention), the big
# stand-alone LayerNorm instances
mkdir -p gpurun_out
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_ffn_tc" -c 8 -f -o gpurun_out/r02_prof_ops2 python scripts/prof_ops.py outproj_ln ffn > gpurun_out/r02_prof_ops2.log 2>&1
tail -2 gpurun_out/r02_prof_ops2.log
if [ "$1" != "quick" ]; then
MLDB_BRANCHES=1 timeout 900 ncu --set full --clock-control none -k regex:"k_ln_vec|k_gemm_tc<256, 1, 2|k_gemm_tc<128" --launch-skip 0 -c 24 -f -o gpurun_out/r02_prof_misc2 python scripts/prof_step.py > gpurun_out/r02_prof_misc2.log 2>&1
tail -2 gpurun_out/r02_prof_misc2.log
fi
