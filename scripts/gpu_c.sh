#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1
tail -4 gpurun_out/pytest_ab.log
for cfg in "MLDB_X=1" "MLDB_FFN_2SM=0"; do
  echo "== $cfg"
  env $cfg timeout 100 python scripts/prof_ops.py qkv outproj_ln ffn 2>&1 | tail -3 | tr '\n' ' '; echo
  env $cfg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_c.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['clocks'])"
done
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_ffn_tc" -c 8 -f -o gpurun_out/prof_ops2 python scripts/prof_ops.py outproj_ln ffn > gpurun_out/prof_ops2.log 2>&1
tail -2 gpurun_out/prof_ops2.log
