"""``B200MLD``: the inference surface of ``mld.models.modeltype.mld.MLD`` (no Lightning).

Methods mirror the reference one to one:
  forward(batch)                      mld.py:216-265
  _diffusion_reverse(emb, lengths)    mld.py:290-360   (one CUDA graph per batch shape)
  gen_from_latent(batch)              mld.py:267-275
  recon_from_motion(batch)            mld.py:277-288
and a scheduler object with the diffusers surface the reference touches
(``init_noise_sigma``, ``set_timesteps``, ``timesteps``, ``step(...).prev_sample``,
``config.num_train_timesteps``; mld.py:310-320,345).

The text encoder is NOT part of this path (frozen CLIP): pass any callable
``text_encoder(List[str]) -> Tensor[2B, S, 768]`` (e.g. the reference's ``MldTextEncoder``).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import torch

from .engine import Engine, make_config


def remove_padding(tensors, lengths):
    """mld/utils/temos_utils.py:24-28."""
    return [t[:n] for t, n in zip(tensors, lengths)]


class B200Scheduler:
    """diffusers ``DDIMScheduler`` / ``DDPMScheduler`` surface backed by the engine's kernels."""

    def __init__(self, engine: Engine):
        self._e = engine
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=engine.cfg.num_train_timesteps)
        self.timesteps: Optional[torch.Tensor] = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.timesteps = self._e.set_timesteps(num_inference_steps)
        return self

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta: float = 0.0, noise=None, **kw):
        if eta != 0.0:
            raise NotImplementedError("only eta == 0 (the shipped config) is built")
        t = int(timestep.reshape(-1)[0]) if torch.is_tensor(timestep) else int(timestep)
        return SimpleNamespace(prev_sample=self._e.scheduler_step(model_output, t, sample, noise))


class B200MLD:
    """Text/action-to-motion sampler with the reference ``MLD`` call surface."""

    def __init__(self, denoiser_sd: Dict[str, torch.Tensor], vae_sd: Dict[str, torch.Tensor], *,
                 mean: torch.Tensor, std: torch.Tensor, text_encoder: Optional[Callable] = None,
                 device: int = 0, num_inference_timesteps: int = 50, condition: str = "text",
                 stage: str = "diffusion", **cfg_kwargs):
        self.cfg = make_config(condition=condition, **cfg_kwargs)
        self.engine = Engine(self.cfg, device)
        self.engine.load_state_dict(denoiser_sd, "denoiser.")
        self.engine.load_state_dict(vae_sd, "vae.")
        self.engine.finalize()
        self.engine.set_mean_std(mean, std)
        self.scheduler = B200Scheduler(self.engine)
        self.scheduler.set_timesteps(num_inference_timesteps)
        self.text_encoder = text_encoder
        self.condition = condition
        self.stage = stage
        self.guidance_scale = float(self.cfg.guidance_scale)
        self.do_classifier_free_guidance = self.guidance_scale > 1.0          # mld.py:115
        self.latent_dim = [self.cfg.n_lat, self.cfg.latent_dim]
        self.device = self.engine.device

    # -- mld.py:216-265 ------------------------------------------------------------------
    def forward(self, batch) -> List[torch.Tensor]:
        lengths = batch["length"]
        if self.stage in ("diffusion", "vae_diffusion"):
            text_emb = self._encode_condition(batch)
            noise = batch.get("init_noise")
            if noise is None:                                                  # mld.py:303-307
                B = len(lengths)
                noise = torch.randn((B, self.latent_dim[0], self.latent_dim[-1]), device=self.device,
                                    dtype=torch.float)
            out = self.engine.sample(text_emb, noise, lengths, want=("joints",))
            joints = out["joints"]
        elif self.stage == "vae":
            z, _ = self._encode_motion(batch["motion"], lengths)
            feats = self.engine.vae_decode(z, lengths)
            joints = self.engine.feats2joints(feats)
        else:
            raise ValueError(self.stage)
        return remove_padding(joints.cpu(), lengths)                          # mld.py:264-265

    __call__ = forward

    def _encode_condition(self, batch) -> torch.Tensor:
        if "text_emb" in batch:                       # pre-computed CLIP output [2B, S, 768]
            return batch["text_emb"]
        if self.condition == "action":
            actions = batch["action"]
            if self.do_classifier_free_guidance:                               # mld.py:716-717
                actions = torch.cat([torch.zeros_like(actions), actions], 0)
            return actions
        texts = list(batch["text"])
        if self.do_classifier_free_guidance:                                   # mld.py:224-230
            texts = [""] * len(texts) + texts
        if self.text_encoder is None:
            raise RuntimeError("no text_encoder was given; pass batch['text_emb'] instead")
        return self.text_encoder(texts)

    # -- mld.py:290-360 ------------------------------------------------------------------
    def _diffusion_reverse(self, encoder_hidden_states: torch.Tensor, lengths=None,
                           init_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        bsz = encoder_hidden_states.shape[0]
        if self.do_classifier_free_guidance:
            bsz = bsz // 2
        if init_noise is None:
            init_noise = torch.randn((bsz, self.latent_dim[0], self.latent_dim[-1]), device=self.device,
                                     dtype=torch.float)
        return self.engine.diffusion_reverse(encoder_hidden_states, init_noise * self.scheduler.init_noise_sigma,
                                             lengths)

    def _encode_motion(self, motion: torch.Tensor, lengths: Sequence[int]):
        mu, logvar = self.engine.vae_encode(motion, lengths)
        std = logvar.exp().pow(0.5)
        dist = torch.distributions.Normal(mu, std)
        return dist.rsample(), dist

    # -- mld.py:267-288 ------------------------------------------------------------------
    def gen_from_latent(self, batch):
        feats = self.engine.vae_decode(batch["latent"], batch["length"])
        return remove_padding(self.engine.feats2joints(feats).cpu(), batch["length"])

    def recon_from_motion(self, batch):
        feats_ref, length = batch["motion"], batch["length"]
        z, _ = self._encode_motion(feats_ref, length)
        feats = self.engine.vae_decode(z, length)
        joints = self.engine.feats2joints(feats).cpu()
        joints_ref = self.engine.feats2joints(feats_ref).cpu()
        return remove_padding(joints, length), remove_padding(joints_ref, length)

    def feats2joints(self, feats: torch.Tensor) -> torch.Tensor:
        return self.engine.feats2joints(feats)
