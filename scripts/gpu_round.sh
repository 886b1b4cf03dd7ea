#!/bin/bash
# one GPU visit: tests, smoke, bench; everything logged under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -15 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -30 gpurun_out/bench.err; cat gpurun_out/bench.log
