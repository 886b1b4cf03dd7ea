#!/bin/bash
# A/B of two library builds inside ONE call (box-to-box clock differences are +-6 %): the tree's libmldb200.so
# against scratch_prev/libmldb200_prev.so (a build of an earlier commit)
mkdir -p gpurun_out
echo "== tests (new build)"
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
cp mld_b200/libmldb200.so /tmp/new.so
for round in 1 2; do for v in new prev; do
  if [ $v = new ]; then cp /tmp/new.so mld_b200/libmldb200.so; else cp scratch_prev/libmldb200_prev.so mld_b200/libmldb200.so; fi
  echo "== $v (round $round)"
  timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5 | tr '\n' ' '; echo
  timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-eager-baseline 2> gpurun_out/bench_ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench', round(d['value'],1), round(d['ms_per_step'],2), d['ddim_step_p50_ms'], d['clocks']['sm_mhz'])" || tail -5 gpurun_out/bench_ab.err
done; done
cp /tmp/new.so mld_b200/libmldb200.so
timeout 200 python scripts/timeline.py outproj_ln 200 > gpurun_out/timeline_ab_outproj.txt 2>&1
