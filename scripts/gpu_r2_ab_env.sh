#!/bin/bash
# A/B of environment knobs inside ONE call.  usage: gpu_r2_ab_env.sh "ENV=a" "ENV=b" ...
mkdir -p gpurun_out
echo "== tests under: $1"
env $1 timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_abe.log 2>&1; tail -3 gpurun_out/pytest_abe.log
for round in 1 2; do for cfg in "$@"; do
  echo "== $cfg (round $round)"
  env $cfg timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo
  env $cfg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_abe.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_abe.err
done; done
