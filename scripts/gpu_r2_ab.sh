#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -q -d POWER | grep -E "Power Draw|Power Limit|Max Power|Min Power" | head -8
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
echo "== operator times (snake on / off)"
for sn in 1 0; do MLDB_SNAKE=$sn timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo; done
echo "== bench"
for cfg in "MLDB_SNAKE=1" "MLDB_SNAKE=0" "MLDB_SNAKE=1 MLDB_BRANCHES=1" "MLDB_SNAKE=0 MLDB_BRANCHES=1"; do env $cfg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_k.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks'])" || tail -5 gpurun_out/bench_k.err; done
echo "== per-op clocks / power (long runs): busiest samples"
for op in ffn qkv attn outproj_ln; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 100 > gpurun_out/smi_k_$op.txt & SMI=$!
  PROF_ITERS=30000 timeout 200 python scripts/prof_ops.py $op 2>&1 | tail -1
  kill $SMI; sort -t, -k2,2nr gpurun_out/smi_k_$op.txt | head -12 | awk -F, '{c+=$1; w+=$2; n++} END{print "   busiest samples:", c/n, "MHz", w/n, "W"}'
done
timeout 200 python scripts/timeline.py attn 120 > gpurun_out/timeline_k_attn.txt 2>&1
