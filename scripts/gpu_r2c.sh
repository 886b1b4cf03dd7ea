#!/bin/bash
# round 2, call C: 8-warp attention groups + FFN tail split + reverted epilogues; first full bench lines
mkdir -p gpurun_out
echo "== kernel unit tests"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/pytest_kernels.log 2>&1; tail -12 gpurun_out/pytest_kernels.log
echo "== parity tests"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_parity.log 2>&1; tail -12 gpurun_out/pytest_parity.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
echo "== operator times"
for cfg in "MLDB_FFN_TAIL=1" "MLDB_FFN_TAIL=0" "MLDB_ATTN=mma"; do echo "$cfg"; env $cfg timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo; done
echo "== bench headline (full line: cpu + eager baselines)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_headline.log 2> gpurun_out/bench_headline.err; tail -3 gpurun_out/bench_headline.err; cut -c1-1800 gpurun_out/bench_headline.log
echo "== bench headline, branches / tail A-B"
for cfg in "MLDB_FFN_TAIL=0" "MLDB_BRANCHES=1" "MLDB_BRANCHES=3"; do env $cfg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_ab.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'])" || tail -3 gpurun_out/bench_ab.err; done
echo "== bench other configs"
for c in 1prompt action512; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err; tail -2 gpurun_out/bench_$c.err; cut -c1-900 gpurun_out/bench_$c.log; done
timeout 900 python bench.py --config novae1024 --steps 1 --warmup 1 > gpurun_out/bench_novae1024.log 2> gpurun_out/bench_novae1024.err; tail -3 gpurun_out/bench_novae1024.err; cut -c1-900 gpurun_out/bench_novae1024.log
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.log
echo "== ncu: hot operators"
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn|k_ffn_tc" -c 20 -f -o gpurun_out/prof_r2c python scripts/prof_ops.py qkv attn outproj_ln ffn > gpurun_out/prof_r2c.log 2>&1; tail -2 gpurun_out/prof_r2c.log
