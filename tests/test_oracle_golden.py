"""CPU: the oracle restatement against fixtures produced by the REFERENCE's own modules
(oracle/make_golden.py).  Tolerances are fp32 re-association noise (the restatement performs
the same fp32 operations; torch's fused MHA fast path may reorder sums)."""
import numpy as np
import torch

from conftest import golden
from mld_b200 import synth
from oracle import mld_oracle as O

torch.set_grad_enabled(False)


def _close(a, b, tol):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert err < tol, f"relative max error {err:.3e} >= {tol}"


def test_denoiser_text_matches_reference():
    g = golden("denoiser_text.npz")
    dsd = synth.denoiser_state_dict(seed=1234)
    for S in (77, 1):
        ctx = synth.text_context(2, S, seed=11)
        x = synth.init_noise(2, seed=12).repeat(2, 1, 1)
        for t in (981, 1):
            y = O.denoiser_forward(dsd, O.DenoiserCfg(), x, torch.tensor(t), ctx, [196, 100] * 2)
            _close(y, g[f"S{S}_t{t}"], 2e-5)


def test_denoiser_action_matches_reference():
    g = golden("denoiser_action.npz")
    asd = synth.denoiser_state_dict(seed=2345, condition="action", num_layers=15, nclasses=12, nfeats=150)
    actions = torch.from_numpy(g["actions"])
    cond = torch.cat([torch.zeros_like(actions), actions])
    x = synth.init_noise(3, seed=22).repeat(2, 1, 1)
    cfg = O.DenoiserCfg(condition="action", num_layers=15, nclasses=12, nfeats=150)
    _close(O.denoiser_forward(asd, cfg, x, torch.tensor(501), cond, [60] * 6), g["y"], 2e-5)


def test_denoiser_novae_matches_reference():
    g = golden("denoiser_novae.npz")
    nsd = synth.denoiser_state_dict(seed=3456, arch="trans_dec", d=512, diffusion_only=True)
    gen = torch.Generator().manual_seed(31)
    x = torch.randn(2, 24, 263, generator=gen).repeat(2, 1, 1)
    ctx = synth.text_context(2, 1, seed=32)
    cfg = O.DenoiserCfg(arch="trans_dec", latent_dim=512, diffusion_only=True)
    y = O.denoiser_forward(nsd, cfg, x, torch.tensor(999), ctx, [24, 16] * 2)
    _close(y, g["y"], 2e-5)
    assert float(y[1, 16:].abs().max()) == 0.0        # padded frames zeroed (mld_denoiser.py:221)


def test_vae_matches_reference():
    g = golden("vae_mld.npz")
    vsd = synth.mld_vae_state_dict(seed=4321)
    lengths = [196, 120, 8]
    z = synth.init_noise(3, seed=41).permute(1, 0, 2).contiguous()
    feats = O.vae_decode(vsd, O.VaeCfg(), z, lengths)
    _close(feats, g["feats"], 2e-5)
    assert float(feats[1, 120:].abs().max()) == 0.0 and float(feats[2, 8:].abs().max()) == 0.0
    gen = torch.Generator().manual_seed(42)
    motion = torch.randn(3, 196, 263, generator=gen)
    mu, logvar = O.vae_encode(vsd, O.VaeCfg(), motion, lengths)
    _close(mu, g["mu"], 2e-5)
    _close(logvar.exp().pow(0.5), g["std"], 2e-5)


def test_actor_vae_matches_reference():
    g = golden("vae_actor.npz")
    avsd = synth.actor_vae_state_dict(seed=777)
    z = synth.init_noise(3, seed=51).permute(1, 0, 2).contiguous()
    feats = O.vae_decode(avsd, O.VaeCfg(kind="actor", nfeats=150, num_layers=6), z, [60, 40, 12])
    _close(feats, g["feats"], 2e-5)


def test_feats2joints_matches_reference():
    g = golden("feats2joints.npz")
    gen = torch.Generator().manual_seed(61)
    f = torch.randn(2, 196, 263, generator=gen) * 0.3
    mean, std = synth.mean_std()
    _close(O.feats2joints(f, mean, std), g["joints"], 1e-6)


def test_sampling_loop_matches_reference_modules():
    """50 guided DDIM steps + decode + feats2joints, S_ctx = 1 (the cheap case on CPU)."""
    g = golden("loop_S1.npz")
    dsd, vsd = synth.denoiser_state_dict(1234), synth.mld_vae_state_dict(4321)
    mean, std = synth.mean_std()
    ctx, noise = synth.text_context(2, 1, seed=71), synth.init_noise(2, seed=72)
    trace = []
    z = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 50, ctx, noise, [196, 88], trace=trace)
    lat = np.stack([t[1].numpy() for t in trace])
    _close(lat, g["latents"], 1e-4)
    feats = O.vae_decode(vsd, O.VaeCfg(), z, [196, 88])
    _close(feats, g["feats"], 1e-4)
    _close(O.feats2joints(feats, mean, std), g["joints"], 1e-4)
