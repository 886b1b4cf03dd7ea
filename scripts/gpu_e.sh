#!/bin/bash
mkdir -p gpurun_out
for cfg in "MLDB_BRANCH_ROUND=1" "MLDB_BRANCH_ROUND=0"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_e.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_e.err | head -3
done
MLDB_BRANCH_ROUND=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
