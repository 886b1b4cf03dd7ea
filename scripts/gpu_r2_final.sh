#!/bin/bash
# round-2 final verification + bench lines for profiles/ (one GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log | cut -c1-400
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err | cut -c1-200; cut -c1-1500 gpurun_out/r02_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-500 gpurun_out/r02_bench_reference_arm.json
timeout 600 python bench.py --config 1prompt --steps 20 --warmup 5 > gpurun_out/r02_bench_1prompt.json 2> gpurun_out/r02_bench_1prompt.err; cut -c1-700 gpurun_out/r02_bench_1prompt.json
timeout 600 python bench.py --config action512 --steps 5 --warmup 3 > gpurun_out/r02_bench_action512.json 2> gpurun_out/r02_bench_action512.err; cut -c1-700 gpurun_out/r02_bench_action512.json
timeout 1200 python bench.py --config novae1024 --steps 1 --warmup 1 > gpurun_out/r02_bench_novae1024.json 2> gpurun_out/r02_bench_novae1024.err; cut -c1-700 gpurun_out/r02_bench_novae1024.json
for op in attn qkv outproj_ln ffn; do timeout 200 python scripts/timeline.py $op 400 > gpurun_out/r02_timeline_$op.txt 2>&1; done
