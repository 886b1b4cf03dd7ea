#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn" -c 25 -f -o gpurun_out/prof_ops python scripts/prof_ops.py > gpurun_out/prof_ops.log 2>&1
tail -15 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/bench.err; cat gpurun_out/bench.log; tail -3 gpurun_out/prof_step.log; tail -8 gpurun_out/prof_ops.log; ls -la gpurun_out
