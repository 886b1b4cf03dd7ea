#!/bin/bash
mkdir -p gpurun_out
echo "== kernels MLDB_TC_2SM=1"
MLDB_TC_2SM=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tc_gemm" > gpurun_out/pytest_2sm.log 2>&1
tail -12 gpurun_out/pytest_2sm.log
if grep -q "passed" gpurun_out/pytest_2sm.log && ! grep -q "failed" gpurun_out/pytest_2sm.log; then
  MLDB_TC_2SM=1 timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_2sm_all.log 2>&1
  tail -4 gpurun_out/pytest_2sm_all.log
  for v in 0 1; do
    echo "== bench MLDB_TC_2SM=$v"
    MLDB_TC_2SM=$v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_2sm$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['op_ms'], d['clocks'])"
    grep -E "Error|error" gpurun_out/bench_2sm$v.err | head -3
  done
fi
