#!/bin/bash
mkdir -p gpurun_out
MLDB_PAIR_CHUNK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "loop or invariance or ragged or vae" > gpurun_out/pytest_pair.log 2>&1
tail -3 gpurun_out/pytest_pair.log
for pc in 1 0; do
  echo "== MLDB_PAIR_CHUNK=$pc"
  MLDB_PAIR_CHUNK=$pc timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_pc$pc.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_pc$pc.err | head -3
done
