#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_br.log 2>&1
tail -5 gpurun_out/pytest_br.log
for b in 1 2 3 4; do
  echo "== bench MLDB_BRANCHES=$b"
  MLDB_BRANCHES=$b timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_br$b.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_br$b.err | head -3
done
