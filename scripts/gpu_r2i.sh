#!/bin/bash
mkdir -p gpurun_out
for sg in 0 2000 6000 12000; do echo "== stagger $sg"; MLDB_FFN_STAGGER=$sg timeout 200 python scripts/prof_ops.py ffn layer 2>&1 | tail -2 | tr '\n' ' '; echo; done
for sg in 0 6000; do MLDB_FFN_STAGGER=$sg timeout 200 python scripts/timeline.py ffn 200 > gpurun_out/timeline_i_ffn_$sg.txt 2>&1; done
timeout 200 python scripts/timeline.py outproj_ln 200 > gpurun_out/timeline_i_outproj_ln.txt 2>&1
echo "== bench"
for sg in 0 6000; do MLDB_FFN_STAGGER=$sg timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_i.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger $sg', round(d['value'],1), round(d['ms_per_step'],2), d['clocks'])" || tail -5 gpurun_out/bench_i.err; done
