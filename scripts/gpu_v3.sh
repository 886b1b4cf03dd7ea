#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_v3.log 2>&1
tail -5 gpurun_out/pytest_v3.log
for fp in 1 0; do
  echo "== MLDB_FFN_PAIR=$fp"
  MLDB_FFN_PAIR=$fp timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_fp$fp.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['op_ms'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_fp$fp.err | head -3
  MLDB_FFN_PAIR=$fp timeout 100 python scripts/prof_ops.py ffn 2>&1 | tail -1
done
