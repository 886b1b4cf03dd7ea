#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/dbg.log
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
bash scripts/gpu_dbg.sh > /dev/null 2>&1
tail -22 gpurun_out/pytest_gpu.log; tail -12 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-400; cat gpurun_out/dbg.log
