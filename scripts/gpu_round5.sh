#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dbg.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc" -c 8 -f -o gpurun_out/prof_ln python scripts/prof_ops.py outproj_ln ffn1 > gpurun_out/prof_ln.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -8 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/prof_ln.log
