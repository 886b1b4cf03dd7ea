#!/bin/bash
mkdir -p gpurun_out
for cl in 2 4; do
  echo "== kernels MLDB_TC_CLUSTER=$cl"
  MLDB_TC_CLUSTER=$cl timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest_cl$cl.log 2>&1
  tail -3 gpurun_out/pytest_cl$cl.log
done
echo "== parity cl=4"
MLDB_TC_CLUSTER=4 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_par_cl4.log 2>&1
tail -3 gpurun_out/pytest_par_cl4.log
for cl in 1 2 4; do
  echo "== bench MLDB_TC_CLUSTER=$cl"
  MLDB_TC_CLUSTER=$cl timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_cl$cl.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['op_ms'], d['clocks'])"
  grep -E "Error|error" gpurun_out/bench_cl$cl.err | head -3
done
