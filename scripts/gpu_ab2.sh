#!/bin/bash
# usage: gpu_ab2.sh "ENV=.. ENV2=.." "ENV=.." ...  (GPU tests once with defaults, then prof_ops + bench per config)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
tail -6 gpurun_out/pytest_ab.log
i=0
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 100 python scripts/prof_ops.py qkv outproj_ln ffn 2>&1 | tail -3 | tr '\n' ' '; echo
  env $cfg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_ab2_$i.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_ab2_$i.err | head -3
  i=$((i+1))
done
