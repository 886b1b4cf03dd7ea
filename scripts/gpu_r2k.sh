#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -q -d POWER | grep -E "Power Draw|Power Limit|Max Power|Min Power" | head -8
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
echo "== bench"
for cfg in "MLDB_BRANCHES=2" "MLDB_BRANCHES=1" "MLDB_BRANCHES=2 MLDB_ATTN=mma"; do env $cfg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_k.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks'])" || tail -5 gpurun_out/bench_k.err; done
echo "== idle-ish op power: prof_ops with many iterations"
(nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 200 > gpurun_out/smi_k.txt &) ; PROF_ITERS=3000 timeout 120 python scripts/prof_ops.py ffn 2>&1 | tail -1; sleep 1; PROF_ITERS=6000 timeout 120 python scripts/prof_ops.py qkv 2>&1 | tail -1;  PROF_ITERS=6000 timeout 120 python scripts/prof_ops.py attn 2>&1 | tail -1; PROF_ITERS=6000 timeout 120 python scripts/prof_ops.py outproj_ln 2>&1 | tail -1
sort gpurun_out/smi_k.txt | uniq -c | sort -k1,1nr | head -12
