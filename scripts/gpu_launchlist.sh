#!/bin/bash
# ncu launch list of the bench command itself (first 700 kernel launches: condition projection + ~6 DDIM steps
# of the first batch, CUDA-graph kernel nodes included); numbers printed by a run under ncu are not bench values
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2> gpurun_out/bench_under_ncu.err
tail -2 gpurun_out/bench_under_ncu.err; wc -l gpurun_out/launches_bench.csv
