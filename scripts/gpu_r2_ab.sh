#!/bin/bash
# A/B template: everything inside ONE call (box-to-box clock differences are +-6 %).  Edit CONFIGS.
mkdir -p gpurun_out
CONFIGS=("MLDB_X=0" "MLDB_FFN_FUSED=0" "MLDB_FFN_SPLIT=0" "MLDB_BRANCHES=1" "MLDB_BRANCHES=3")
for cfg in "${CONFIGS[@]}"; do
  echo "== $cfg"
  env $cfg timeout 200 python scripts/prof_ops.py ffn ffn1 ffn2_ln layer 2>&1 | tail -4 | tr '\n' ' '; echo
  for rep in 1; do env $cfg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_ab.err; done
done
env ${CONFIGS[0]} timeout 200 python scripts/timeline.py attn 120 > gpurun_out/timeline_ab_attn.txt 2>&1
