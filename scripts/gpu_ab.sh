#!/bin/bash
# usage: gpu_ab.sh VAR v1 v2 ...   (runs the GPU tests once, then the bench for each value of VAR)
mkdir -p gpurun_out
VAR=$1; shift
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
tail -4 gpurun_out/pytest_ab.log
for v in "$@"; do
  echo "== bench $VAR=$v"
  env $VAR=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_ab_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['op_ms'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_ab_$v.err | head -3
  env $VAR=$v timeout 100 python scripts/prof_ops.py ffn 2>&1 | tail -1
done
