#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_ffn" > gpurun_out/pytest_ffn.log 2>&1
tail -15 gpurun_out/pytest_ffn.log
if grep -q "passed" gpurun_out/pytest_ffn.log && ! grep -q "failed" gpurun_out/pytest_ffn.log; then
  timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_all.log 2>&1
  tail -5 gpurun_out/pytest_all.log
  for f in 1 0; do
    echo "== bench MLDB_FFN_FUSED=$f"
    MLDB_FFN_FUSED=$f timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_ff$f.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['op_ms'], d['clocks'])"
    grep -E "Error|error" gpurun_out/bench_ff$f.err | head -3
    MLDB_FFN_FUSED=$f timeout 100 python scripts/prof_ops.py ffn 2>&1 | tail -1
  done
fi
