#!/bin/bash
# round 2, call A: bring-up of the tcgen05 attention kernel + the rewritten LayerNorm epilogues.
# Every group in its own process (a device trap poisons a context); everything bounded by timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
echo "== attention bring-up (per-shape processes)"
timeout 900 python scripts/dbg_attn.py > gpurun_out/dbg_attn.log 2>&1; tail -60 gpurun_out/dbg_attn.log
echo "== kernel unit tests: gemm / ffn"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm or ffn_block" > gpurun_out/pytest_gemm.log 2>&1; tail -15 gpurun_out/pytest_gemm.log
echo "== kernel unit tests: attention"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" > gpurun_out/pytest_attn.log 2>&1; tail -15 gpurun_out/pytest_attn.log
echo "== kernel unit tests: whole path / options / stats"
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "whole_path or scheduling or product_path" > gpurun_out/pytest_path.log 2>&1; tail -15 gpurun_out/pytest_path.log
echo "== parity tests (attn=tc default)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_parity.log 2>&1; tail -25 gpurun_out/pytest_parity.log
echo "== parity tests with the mma.sync attention (isolates attention from the GEMM changes)"
MLDB_ATTN=mma timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not novae_full and not two_latent" > gpurun_out/pytest_parity_mma.log 2>&1; tail -8 gpurun_out/pytest_parity_mma.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
echo "== operator times"
for a in tc mma; do echo "attn=$a"; MLDB_ATTN=$a timeout 200 python scripts/prof_ops.py qkv attn outproj_ln ffn layer 2>&1 | tail -5; done
echo "== bench"
for a in tc mma; do MLDB_ATTN=$a timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$a.log 2> gpurun_out/bench_$a.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$a.log").read().strip().splitlines()[-1])
    print("attn=$a", round(d["value"], 1), "motions/s", round(d["ms_per_step"], 2), "ms/batch e2e", round(d["e2e"]["value"], 1), d["clocks"], d["roofline"]["op_ms"], d["roofline"]["layer_ms"])
except Exception as e:
    print("attn=$a bench failed:", e); print(open("gpurun_out/bench_$a.err").read()[-1500:])
PY
done
echo "== ncu: hot operators"
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn|k_ffn_tc" -c 16 -f -o gpurun_out/prof_r2b python scripts/prof_ops.py qkv attn outproj_ln ffn > gpurun_out/prof_r2b.log 2>&1; tail -3 gpurun_out/prof_r2b.log
ls -la gpurun_out | tail -5
