#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
for pdl in 1 0; do
  echo "== MLDB_PDL=$pdl"
  MLDB_PDL=$pdl timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_pdl$pdl.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['op_ms'], d['roofline']['layer_ms'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_pdl$pdl.err | head -3
done
