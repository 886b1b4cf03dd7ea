#!/bin/bash
mkdir -p gpurun_out
echo "== tcgen05.mma microbenchmark"
timeout 120 scripts/ubench/mma_lat > gpurun_out/mma_lat.txt 2>&1; cat gpurun_out/mma_lat.txt
echo "== tests (res_mma default on)"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_d.log 2>&1; tail -8 gpurun_out/pytest_d.log
echo "== operator times"
for cfg in "MLDB_RES_MMA=1 MLDB_FFN_TAIL=0" "MLDB_RES_MMA=0 MLDB_FFN_TAIL=0"; do echo "$cfg"; env $cfg timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo; done
echo "== bench A/B"
for cfg in "MLDB_RES_MMA=1 MLDB_FFN_TAIL=0" "MLDB_RES_MMA=0 MLDB_FFN_TAIL=0" "MLDB_RES_MMA=1 MLDB_FFN_TAIL=0 MLDB_BRANCHES=1"; do env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks']['sm_mhz'])" || tail -3 gpurun_out/bench_ab.err; done
