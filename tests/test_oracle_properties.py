"""CPU: size-independent properties of the restated sampling path (oracle = test infrastructure).
They hold for the reference by construction and are what the GPU parity tests lean on at sizes the
oracle cannot reach: motions are independent of their batch mates, guidance with scale 1 leaves the
conditional prediction, padded frames decode to exact zeros, and feats2joints is equivariant to the
root's starting pose (a pure prefix sum)."""
import torch

from mld_b200 import synth
from oracle import mld_oracle as O

torch.set_grad_enabled(False)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_motions_are_independent_of_batch_mates():
    dsd = synth.denoiser_state_dict(seed=1234)
    ctx = synth.text_context(3, 5, seed=3)            # [2B, S, 768], uncond half first
    noise = synth.init_noise(3, seed=4)               # [B, 1, 256]
    full = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 3, ctx, noise)
    for b in range(3):
        sel = torch.tensor([b, 3 + b])
        one = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 3, ctx[sel], noise[b:b + 1])
        assert _rel(one[:, 0], full[:, b]) < 1e-5


def test_guidance_scale_one_is_the_conditional_path():
    """mld.py:325-342: with guidance_scale <= 1 the batch is not doubled and eps is the conditional eps."""
    dsd = synth.denoiser_state_dict(seed=1234)
    ctx = synth.text_context(2, 4, seed=5)
    noise = synth.init_noise(2, seed=6)
    cond_only = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 2, ctx[2:], noise, guidance_scale=1.0)
    # u + 1.0000001 * (c - u) ~= c : guided path with a scale barely above 1
    guided = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 2, ctx, noise, guidance_scale=1.0 + 1e-7)
    assert _rel(guided, cond_only) < 1e-5


def test_padded_frames_decode_to_exact_zero_and_do_not_leak():
    vsd = synth.mld_vae_state_dict(seed=4321)
    z = torch.randn(1, 2, 256, generator=torch.Generator().manual_seed(1))
    feats = O.vae_decode(vsd, O.VaeCfg(), z, [40, 64])
    assert feats.shape[1] == 64
    assert torch.count_nonzero(feats[0, 40:]) == 0                       # mld_vae.py:245
    # the shorter motion decoded alone (its own max length) gives the same valid frames
    alone = O.vae_decode(vsd, O.VaeCfg(), z[:, :1], [40])
    assert _rel(alone[0, :40], feats[0, :40]) < 1e-5


def test_ddim_loop_is_deterministic_and_timestep_indexing_is_integer_exact():
    sch = O.DDIMScheduler()
    sch.set_timesteps(50)
    ts = [int(t) for t in sch.timesteps]
    assert ts[0] == 981 and ts[-1] == 1 and all(a - b == 20 for a, b in zip(ts, ts[1:]))   # steps_offset = 1
    dsd = synth.denoiser_state_dict(seed=1234)
    ctx, noise = synth.text_context(1, 3, seed=7), synth.init_noise(1, seed=8)
    a = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 4, ctx, noise)
    b = O.diffusion_reverse(dsd, O.DenoiserCfg(), O.DDIMScheduler(), 4, ctx, noise)
    assert torch.equal(a, b)


def test_feats2joints_is_a_prefix_sum_of_root_motion():
    """recover_from_ric integrates the root's angular and planar velocities (motion_process.py:362-381):
    the joints of a motion's prefix do not depend on later frames."""
    mean, std = synth.mean_std()
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(1, 32, 263, generator=g)
    full = O.feats2joints(feats, mean, std)
    head = O.feats2joints(feats[:, :20], mean, std)
    assert _rel(head, full[:, :20]) < 1e-6
    assert full.shape == (1, 32, 22, 3)
