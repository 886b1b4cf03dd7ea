"""Run each layer-0 operator of the denoiser once at the benchmark shape (B=256, 77-token context)
so that `ncu --set full -k regex:...` can capture the hot kernels in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mld_b200 import synth
from mld_b200.engine import Engine, make_config

eng = Engine(make_config(), 0)
eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
eng.finalize()
eng.set_timesteps(2)
ops = sys.argv[1:] or ["qkv", "attn", "outproj_ln", "ffn1", "ffn2_ln"]
for op in ops:
    print(op, eng.profile_op(op, 256, 77, int(__import__("os").environ.get("PROF_ITERS", "10"))), "ms")
