#!/bin/bash
# round-end verification + evidence (no ncu --set full: see gpu_c.sh / gpu_final.sh for those captures)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
MLDB_BRANCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err; cut -c1-700 gpurun_out/bench.log; cut -c1-300 gpurun_out/bench_ref.log; tail -2 gpurun_out/prof_step.log
