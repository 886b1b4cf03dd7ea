#!/bin/bash
# round-end evidence: launch list (kernel shares of two DDIM steps + decode), ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py > gpurun_out/prof_step.log 2>&1
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn" -c 20 -f -o gpurun_out/prof_ops python scripts/prof_ops.py > gpurun_out/prof_ops.log 2>&1
timeout 300 python - <<'PY' > gpurun_out/decode_time.log 2>&1
import sys, torch
sys.path.insert(0, '.')
from mld_b200 import synth
from mld_b200.engine import Engine, make_config
eng = Engine(make_config(), 0)
eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser."); eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
eng.finalize(); eng.set_mean_std(*synth.mean_std()); eng.set_timesteps(50)
B = 256
z = synth.init_noise(B, seed=3).permute(1, 0, 2).contiguous().cuda()
lengths = [196] * B
for name, fn in (("decode", lambda: eng.vae_decode(z, lengths)),):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms")
f = eng.vae_decode(z, lengths)
for _ in range(3): eng.feats2joints(f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.feats2joints(f)
e1.record(); torch.cuda.synchronize()
print("feats2joints", e0.elapsed_time(e1) / 10, "ms")
PY
tail -2 gpurun_out/prof_step.log; tail -3 gpurun_out/prof_ops.log; cat gpurun_out/decode_time.log
