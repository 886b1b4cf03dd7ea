#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
tail -4 gpurun_out/pytest_ab.log
for cfg in "MLDB_X=1" "MLDB_LANES=0" "MLDB_BRANCHES=3" "MLDB_BRANCHES=4"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_d.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_d.err | head -3
done
