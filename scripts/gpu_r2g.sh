#!/bin/bash
mkdir -p gpurun_out
echo "== attention bring-up (per-shape processes, short timeouts)"
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_h.log 2>&1; tail -8 gpurun_out/pytest_g.log
echo "== operator times"
timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo
echo "== bench"
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks'], d['e2e']['value'], d.get('gpu_eager_baseline'))" || tail -5 gpurun_out/bench_h.err
for op in attn qkv outproj_ln ffn; do echo "== timeline $op"; timeout 200 python scripts/timeline.py $op 120 > gpurun_out/timeline_h_$op.txt 2>&1; head -40 gpurun_out/timeline_g_$op.txt | cut -c1-260; done
