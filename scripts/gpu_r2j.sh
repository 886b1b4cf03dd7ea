#!/bin/bash
mkdir -p gpurun_out
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_j.log 2>&1; tail -8 gpurun_out/pytest_j.log
echo "== operator times"
for sp in 0 1; do echo "split $sp"; MLDB_FFN_SPLIT=$sp timeout 200 python scripts/prof_ops.py ffn layer 2>&1 | tail -2 | tr '\n' ' '; echo; done
echo "== bench"
for cfg in "MLDB_FFN_SPLIT=0 MLDB_BRANCHES=2" "MLDB_FFN_SPLIT=1 MLDB_BRANCHES=2" "MLDB_FFN_SPLIT=1 MLDB_BRANCHES=1" "MLDB_FFN_SPLIT=0 MLDB_BRANCHES=1"; do env $cfg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_j.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_j.err; done
for sp in 0 1; do MLDB_FFN_SPLIT=$sp timeout 300 python bench.py --config 1prompt --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_j1.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1prompt split $sp', round(d['value'],2), round(d['ms_per_step'],3), d['ddim_step_p50_ms'])" || tail -5 gpurun_out/bench_j1.err; done
MLDB_FFN_SPLIT=1 timeout 200 python scripts/timeline.py ffn 200 > gpurun_out/timeline_j_ffn.txt 2>&1
