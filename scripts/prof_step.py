"""Two DDIM steps + decode + joints at the benchmark shape with graphs off, for the ncu launch list
(`ncu --metrics gpu__time_duration.sum`): per-kernel share of a step."""
import os, sys
os.environ["MLDB_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mld_b200 import synth
from mld_b200.engine import Engine, make_config

B, S = 256, 77
eng = Engine(make_config(), 0)
eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
eng.finalize()
eng.set_mean_std(*synth.mean_std())
eng.set_timesteps(2)
ctx, noise = synth.text_context(B, S, seed=1).cuda(), synth.init_noise(B, seed=2).cuda()
out = eng.sample(ctx, noise, [196] * B, want=("joints",))
torch.cuda.synchronize()
print("launches", eng.launch_count)
