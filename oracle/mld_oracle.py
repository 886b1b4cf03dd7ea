"""fp32 CPU restatement of the reference's sampling path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference file:line it follows (paths relative to the
ChenFengYe/motion-latent-diffusion tree).  Weights are plain ``dict[str, Tensor]`` with the
reference's own state-dict key names; tensors use the reference's sequence-first layout
``[L, B, d]`` so the code reads like the code it mirrors.

Parity: pinned against the reference modules through ``tests/golden`` (see
``oracle/make_golden.py``).  Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU
baseline legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- primitives
def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    """nn.Linear: y = x W^T + b."""
    return F.linear(x, w, b)


def layer_norm(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """nn.LayerNorm(d), eps 1e-5 (torch default; cross_attention.py:247-248)."""
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + "weight"], sd[prefix + "bias"], 1e-5)


def lengths_to_mask(lengths: Sequence[int], max_len: Optional[int] = None) -> Tensor:
    """mld/utils/temos_utils.py:10-17 -> bool[B, max_len], True = valid frame."""
    lengths_t = torch.as_tensor(list(lengths), dtype=torch.long)
    max_len = int(max_len) if max_len else int(lengths_t.max())
    return torch.arange(max_len).expand(len(lengths_t), max_len) < lengths_t.unsqueeze(1)


def timestep_features(timesteps: Tensor, dim: int, flip_sin_to_cos: bool = True,
                      freq_shift: float = 0.0, max_period: int = 10000) -> Tensor:
    """mld/models/architectures/tools/embeddings.py:245-285 (get_timestep_embedding)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


def time_mlp(sd: SD, prefix: str, feats: Tensor) -> Tensor:
    """TimestepEmbedding: Linear -> SiLU -> Linear (embeddings.py:288-305)."""
    h = linear(feats, sd[prefix + "linear_1.weight"], sd[prefix + "linear_1.bias"])
    h = F.silu(h)
    return linear(h, sd[prefix + "linear_2.weight"], sd[prefix + "linear_2.bias"])


def mha(query: Tensor, key: Tensor, value: Tensor, sd: SD, prefix: str, nhead: int,
        key_padding_mask: Optional[Tensor] = None) -> Tensor:
    """nn.MultiheadAttention forward (eval, no dropout), as used at
    cross_attention.py:264-266,330-338.  Packed in_proj rows are [Wq; Wk; Wv]; heads are
    contiguous d/nhead slices; scale 1/sqrt(head_dim); masked keys -> -inf.
    query [L,B,d], key/value [S,B,d], key_padding_mask bool[B,S] (True = ignore)."""
    L, B, d = query.shape
    S = key.shape[0]
    hd = d // nhead
    w, b = sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"]
    q = linear(query, w[:d], b[:d])
    k = linear(key, w[d:2 * d], b[d:2 * d])
    v = linear(value, w[2 * d:], b[2 * d:])
    q = q.reshape(L, B * nhead, hd).transpose(0, 1)          # [B*h, L, hd]
    k = k.reshape(S, B * nhead, hd).transpose(0, 1)
    v = v.reshape(S, B * nhead, hd).transpose(0, 1)
    scores = torch.bmm(q * (1.0 / math.sqrt(hd)), k.transpose(1, 2))   # [B*h, L, S]
    if key_padding_mask is not None:
        m = key_padding_mask.view(B, 1, 1, S).expand(B, nhead, 1, S).reshape(B * nhead, 1, S)
        scores = scores.masked_fill(m, float("-inf"))
    p = torch.softmax(scores, dim=-1)
    o = torch.bmm(p, v)                                      # [B*h, L, hd]
    o = o.transpose(0, 1).reshape(L, B, d)
    return linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def _act(name: str):
    """cross_attention.py:404-412 (_get_activation_fn); gelu is the exact erf form."""
    return {"gelu": F.gelu, "relu": F.relu}[name]


def encoder_layer_post(x: Tensor, sd: SD, p: str, nhead: int,
                       kpm: Optional[Tensor] = None, act: str = "gelu") -> Tensor:
    """TransformerEncoderLayer.forward_post (cross_attention.py:259-272)."""
    x = layer_norm(x + mha(x, x, x, sd, p + "self_attn.", nhead, kpm), sd, p + "norm1.")
    h = _act(act)(linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
    h = linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(x + h, sd, p + "norm2.")


def decoder_layer_post(tgt: Tensor, memory: Tensor, sd: SD, p: str, nhead: int,
                       tgt_kpm: Optional[Tensor] = None, act: str = "gelu") -> Tensor:
    """TransformerDecoderLayer.forward_post (cross_attention.py:323-345); the same math as
    torch's nn.TransformerDecoderLayer (norm_first=False) used by ActorVae
    (actor_vae.py:195-203).  No memory mask on either call site."""
    tgt = layer_norm(tgt + mha(tgt, tgt, tgt, sd, p + "self_attn.", nhead, tgt_kpm), sd, p + "norm1.")
    tgt = layer_norm(tgt + mha(tgt, memory, memory, sd, p + "multihead_attn.", nhead), sd, p + "norm2.")
    h = _act(act)(linear(tgt, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
    h = linear(h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return layer_norm(tgt + h, sd, p + "norm3.")


def skip_encoder(x: Tensor, sd: SD, p: str, num_layers: int, nhead: int,
                 kpm: Optional[Tensor] = None, act: str = "gelu") -> Tensor:
    """SkipTransformerEncoder.forward (cross_attention.py:41-64): (L-1)/2 input blocks
    pushed on a LIFO stack, a middle block, (L-1)/2 output blocks each preceded by
    Linear(2d->d)(cat[x, stack.pop()]), final LayerNorm."""
    nb = (num_layers - 1) // 2
    xs = []
    for i in range(nb):
        x = encoder_layer_post(x, sd, f"{p}input_blocks.{i}.", nhead, kpm, act)
        xs.append(x)
    x = encoder_layer_post(x, sd, f"{p}middle_block.", nhead, kpm, act)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = linear(x, sd[f"{p}linear_blocks.{i}.weight"], sd[f"{p}linear_blocks.{i}.bias"])
        x = encoder_layer_post(x, sd, f"{p}output_blocks.{i}.", nhead, kpm, act)
    return layer_norm(x, sd, p + "norm.")


def skip_decoder(tgt: Tensor, memory: Tensor, sd: SD, p: str, num_layers: int, nhead: int,
                 tgt_kpm: Optional[Tensor] = None, act: str = "gelu") -> Tensor:
    """SkipTransformerDecoder.forward (cross_attention.py:89-125)."""
    nb = (num_layers - 1) // 2
    x, xs = tgt, []
    for i in range(nb):
        x = decoder_layer_post(x, memory, sd, f"{p}input_blocks.{i}.", nhead, tgt_kpm, act)
        xs.append(x)
    x = decoder_layer_post(x, memory, sd, f"{p}middle_block.", nhead, tgt_kpm, act)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = linear(x, sd[f"{p}linear_blocks.{i}.weight"], sd[f"{p}linear_blocks.{i}.bias"])
        x = decoder_layer_post(x, memory, sd, f"{p}output_blocks.{i}.", nhead, tgt_kpm, act)
    return layer_norm(x, sd, p + "norm.")


def plain_decoder(tgt: Tensor, memory: Tensor, sd: SD, p: str, num_layers: int, nhead: int,
                  tgt_kpm: Optional[Tensor] = None, act: str = "gelu",
                  final_norm: bool = True) -> Tensor:
    """TransformerDecoder.forward (cross_attention.py:204-233, layers + decoder_norm) and
    torch's nn.TransformerDecoder without a final norm (actor_vae.py:205-206)."""
    x = tgt
    for i in range(num_layers):
        x = decoder_layer_post(x, memory, sd, f"{p}layers.{i}.", nhead, tgt_kpm, act)
    return layer_norm(x, sd, p + "norm.") if final_norm else x


def sine_pe(n: int, d: int) -> Tensor:
    """PositionalEncoding buffer rows [n, d] (position_encoding_layer.py:14-20)."""
    pe = torch.zeros(n, d)
    position = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-np.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


# --------------------------------------------------------------------------- configs
@dataclass
class DenoiserCfg:
    """ctor kwargs of MldDenoiser that change the math (mld_denoiser.py:18-38)."""
    condition: str = "text"            # "text" | "action"
    arch: str = "trans_enc"            # "trans_enc" (skip) | "trans_dec" (no-VAE)
    latent_dim: int = 256
    n_lat: int = 1
    ff_size: int = 1024
    num_layers: int = 9
    num_heads: int = 4
    text_encoded_dim: int = 768
    flip_sin_to_cos: bool = True
    freq_shift: float = 0.0
    activation: str = "gelu"
    diffusion_only: bool = False       # ablation.VAE_TYPE == "no"
    nfeats: int = 263
    nclasses: int = 12
    guidance_scale: float = 7.5


@dataclass
class VaeCfg:
    """ctor kwargs of MldVae / ActorVae (mld_vae.py:35-47, actor_vae.py:13-24)."""
    kind: str = "mld"                  # "mld" (encoder_decoder, learned PE) | "actor"
    nfeats: int = 263
    latent_dim: int = 256
    n_lat: int = 1
    ff_size: int = 1024
    num_layers: int = 9
    num_heads: int = 4
    activation: str = "gelu"


# --------------------------------------------------------------------------- denoiser
def embed_action(sd: SD, actions: Tensor, guidance_scale: float) -> Tensor:
    """EmbedAction.forward in eval mode (mld_denoiser.py:250-262): gather rows, force the
    first (uncond) half of the CFG batch to zeros, unsqueeze(0)."""
    idx = actions[:, 0].to(torch.long)
    out = sd["emb_proj.action_embedding"][idx]
    if guidance_scale > 1.0:
        uncond, cond = out.chunk(2)
        out = torch.cat((torch.zeros_like(uncond), cond))
    return out.unsqueeze(0)


def denoiser_forward(sd: SD, cfg: DenoiserCfg, sample: Tensor, timestep: Tensor,
                     encoder_hidden_states: Tensor,
                     lengths: Optional[Sequence[int]] = None) -> Tensor:
    """MldDenoiser.forward (mld_denoiser.py:135-228).  sample [Bx, n_lat, d] (or
    [Bx, T, nfeats] when diffusion_only); returns the same shape (the reference wraps it in
    a 1-tuple)."""
    d = cfg.latent_dim
    sample = sample.permute(1, 0, 2)                                  # :143
    Bx = sample.shape[1]
    mask = lengths_to_mask(lengths) if lengths not in (None, []) else None   # :146-147
    timesteps = timestep.expand(Bx).clone()                          # :151
    tdim = cfg.text_encoded_dim if cfg.condition == "text" else d     # :57,70
    time_emb = timestep_features(timesteps, tdim, cfg.flip_sin_to_cos, cfg.freq_shift)
    time_emb = time_mlp(sd, "time_embedding.", time_emb).unsqueeze(0)  # :155

    if cfg.condition == "text":
        text_emb = encoder_hidden_states.permute(1, 0, 2)             # :162
        if cfg.text_encoded_dim != d:                                 # :165 ReLU -> Linear
            text_emb = linear(F.relu(text_emb), sd["emb_proj.1.weight"], sd["emb_proj.1.bias"])
        emb_latent = torch.cat((time_emb, text_emb), 0)               # :171
    elif cfg.condition == "action":
        emb_latent = torch.cat((time_emb, embed_action(sd, encoder_hidden_states,
                                                       cfg.guidance_scale)), 0)   # :173-177
    else:
        raise TypeError(f"condition type {cfg.condition} not supported")

    if cfg.arch == "trans_enc":
        if cfg.diffusion_only:
            sample = linear(sample, sd["pose_embd.weight"], sd["pose_embd.bias"])
            xseq = torch.cat((emb_latent, sample), 0)                 # :184-185
        else:
            xseq = torch.cat((sample, emb_latent), 0)                 # :187
        xseq = xseq + sd["query_pos.pe"][: xseq.shape[0]]             # :196, position_encoding.py:158
        tokens = skip_encoder(xseq, sd, "encoder.", cfg.num_layers, cfg.num_heads, None,
                              cfg.activation)                          # :197
        if cfg.diffusion_only:
            out = tokens[emb_latent.shape[0]:]
            out = linear(out, sd["pose_proj.weight"], sd["pose_proj.bias"])
            out = out.clone()
            out[~mask.T] = 0                                          # :204
        else:
            out = tokens[: sample.shape[0]]                           # :206
    elif cfg.arch == "trans_dec":
        if cfg.diffusion_only:
            sample = linear(sample, sd["pose_embd.weight"], sd["pose_embd.bias"])   # :210
        sample = sample + sd["query_pos.pe"][: sample.shape[0]]       # :214
        emb_latent = emb_latent + sd["mem_pos.pe"][: emb_latent.shape[0]]   # :215
        out = plain_decoder(sample, emb_latent, sd, "decoder.", cfg.num_layers, cfg.num_heads,
                            None, cfg.activation, final_norm=True)     # :216 (unsqueeze/squeeze)
        if cfg.diffusion_only:
            out = linear(out, sd["pose_proj.weight"], sd["pose_proj.bias"]).clone()
            out[~mask.T] = 0                                          # :219-221
    else:
        raise TypeError(f"{cfg.arch} is not supported")
    return out.permute(1, 0, 2)                                       # :226


# --------------------------------------------------------------------------- VAE
def vae_decode(sd: SD, cfg: VaeCfg, z: Tensor, lengths: Sequence[int]) -> Tensor:
    """MldVae.decode, arch encoder_decoder, pe_type mld (mld_vae.py:186-248) and
    ActorAgnosticDecoder.forward (actor_vae.py:210-235).  z [n_lat,B,d] -> feats [B,T,F],
    padded frames exactly zero."""
    mask = lengths_to_mask(lengths)
    B, T = mask.shape
    d = cfg.latent_dim
    queries = torch.zeros(T, B, d)
    if cfg.kind == "mld":
        queries = queries + sd["query_pos_decoder.pe"][:T]            # :224
        out = skip_decoder(queries, z, sd, "decoder.", cfg.num_layers, cfg.num_heads, ~mask,
                           cfg.activation)                             # :226-232
        out = linear(out, sd["final_layer.weight"], sd["final_layer.bias"])   # :243
    elif cfg.kind == "actor":
        queries = queries + sine_pe(T, d).unsqueeze(1)                 # actor_vae.py:221-225
        out = plain_decoder(queries, z, sd, "decoder.seqTransDecoder.", cfg.num_layers,
                            cfg.num_heads, ~mask, cfg.activation, final_norm=False)
        out = linear(out, sd["decoder.final_layer.weight"], sd["decoder.final_layer.bias"])
    else:
        raise ValueError(cfg.kind)
    out = out.clone()
    out[~mask.T] = 0                                                   # :245
    return out.permute(1, 0, 2)                                        # :247


def vae_encode(sd: SD, cfg: VaeCfg, feats: Tensor, lengths: Sequence[int]):
    """MldVae.encode up to (mu, logvar) (mld_vae.py:124-178); the rsample() at :181-183 is
    torch RNG and is left to the caller (z = mu + exp(0.5 logvar) * eps)."""
    B, T, _ = feats.shape
    mask = lengths_to_mask(lengths, T)
    x = linear(feats, sd["skel_embedding.weight"], sd["skel_embedding.bias"]).permute(1, 0, 2)
    dist = torch.tile(sd["global_motion_token"][:, None, :], (1, B, 1))   # :146
    aug = torch.cat((torch.ones(B, dist.shape[0], dtype=torch.bool), mask), 1)
    xseq = torch.cat((dist, x), 0)
    xseq = xseq + sd["query_pos_encoder.pe"][: xseq.shape[0]]          # :161
    out = skip_encoder(xseq, sd, "encoder.", cfg.num_layers, cfg.num_heads, ~aug,
                       cfg.activation)[: dist.shape[0]]
    return out[: cfg.n_lat], out[cfg.n_lat:]                           # mu, logvar :177-178


# --------------------------------------------------------------------------- schedulers
class DDIMScheduler:
    """diffusers.DDIMScheduler restated (diffusers is NOT in the reference tree; unpinned in
    requirements.txt:23).  Config per configs/modules/scheduler.yaml:1-14; call sites
    mld.py:81,310-320,345.  prediction_type epsilon, eta given at step()."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                 steps_offset=1):
        assert beta_schedule == "scaled_linear" and not clip_sample
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.steps_offset = steps_offset
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset

    def step(self, model_output: Tensor, timestep, sample: Tensor, eta: float = 0.0,
             noise: Optional[Tensor] = None) -> Tensor:
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - a_t
        pred_x0 = (sample - beta_prod_t ** 0.5 * model_output) / a_t ** 0.5
        variance = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        std_dev_t = eta * variance ** 0.5
        direction = (1 - a_prev - std_dev_t ** 2) ** 0.5 * model_output
        prev = a_prev ** 0.5 * pred_x0 + direction
        if eta > 0:
            prev = prev + std_dev_t * noise
        return prev


class DDPMScheduler:
    """diffusers.DDPMScheduler restated (configs/modules_novae/scheduler.yaml:16-29:
    fixed_small variance, no clipping, epsilon prediction).  The per-step noise is injected
    by the caller (diffusers draws it inside step(); device/order is version dependent)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", variance_type="fixed_small", clip_sample=False):
        assert beta_schedule == "scaled_linear" and variance_type == "fixed_small" and not clip_sample
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts)

    def step(self, model_output: Tensor, timestep, sample: Tensor,
             noise: Optional[Tensor] = None) -> Tensor:
        t = int(timestep)
        n = self.num_inference_steps or self.num_train_timesteps
        prev_t = t - self.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - a_t
        beta_prod_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        pred_x0 = (sample - beta_prod_t ** 0.5 * model_output) / a_t ** 0.5
        c0 = (a_prev ** 0.5 * cur_beta) / beta_prod_t
        c1 = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        prev = c0 * pred_x0 + c1 * sample
        if t > 0:
            var = torch.clamp(beta_prod_prev / beta_prod_t * cur_beta, min=1e-20)
            prev = prev + var ** 0.5 * noise
        return prev

    def add_noise(self, x0: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
        a = self.alphas_cumprod[timesteps]
        sa = (a ** 0.5).flatten()
        sb = ((1 - a) ** 0.5).flatten()
        while sa.dim() < x0.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * x0 + sb * noise


# --------------------------------------------------------------------------- sampler loop
def diffusion_reverse(dsd: SD, dcfg: DenoiserCfg, scheduler, num_inference_steps: int,
                      encoder_hidden_states: Tensor, init_noise: Tensor,
                      lengths: Optional[Sequence[int]] = None, guidance_scale: float = 7.5,
                      eta: float = 0.0, step_noise: Optional[Tensor] = None,
                      trace: Optional[list] = None) -> Tensor:
    """MLD._diffusion_reverse (mld.py:290-360) with the initial latents passed in instead of
    drawn (:303-307).  encoder_hidden_states is [2B,...] with the uncond half first
    (:225-230).  Returns [n_lat, B, d] (:359), or [T, B, F] for the no-VAE model."""
    cfg_on = guidance_scale > 1.0
    latents = init_noise * scheduler.init_noise_sigma                 # :310
    scheduler.set_timesteps(num_inference_steps)                      # :312
    is_ddim = isinstance(scheduler, DDIMScheduler)
    for i, t in enumerate(scheduler.timesteps.to(latents.device)):    # :323 (from_numpy ignores a device context)
        model_in = torch.cat([latents] * 2) if cfg_on else latents    # :325-327
        lengths_rev = (list(lengths) * 2 if cfg_on else lengths) if lengths is not None else None
        noise_pred = denoiser_forward(dsd, dcfg, model_in, t, encoder_hidden_states, lengths_rev)
        if cfg_on:                                                    # :339-342
            u, c = noise_pred.chunk(2)
            noise_pred = u + guidance_scale * (c - u)
        if is_ddim:
            latents = scheduler.step(noise_pred, t, latents, eta=eta)  # :345
        else:
            latents = scheduler.step(noise_pred, t, latents,
                                     noise=None if step_noise is None else step_noise[i])
        if trace is not None:
            trace.append((noise_pred.clone(), latents.clone()))
    return latents.permute(1, 0, 2)                                   # :359


# --------------------------------------------------------------------------- feats2joints
def qinv(q: Tensor) -> Tensor:
    """mld/data/humanml/common/quaternion.py:16-20."""
    mask = torch.ones_like(q)
    mask[..., 1:] = -mask[..., 1:]
    return q * mask


def qrot(q: Tensor, v: Tensor) -> Tensor:
    """mld/data/humanml/common/quaternion.py:54-73."""
    shape = list(v.shape)
    q = q.contiguous().view(-1, 4)
    v = v.contiguous().view(-1, 3)
    qvec = q[:, 1:]
    uv = torch.cross(qvec, v, dim=1)
    uuv = torch.cross(qvec, uv, dim=1)
    return (v + 2 * (q[:, :1] * uv + uuv)).view(shape)


def recover_root_rot_pos(data: Tensor):
    """mld/data/humanml/scripts/motion_process.py:362-381."""
    rot_vel = data[..., 0]
    r_rot_ang = torch.zeros_like(rot_vel)
    r_rot_ang[..., 1:] = rot_vel[..., :-1]
    r_rot_ang = torch.cumsum(r_rot_ang, dim=-1)
    r_rot_quat = torch.zeros(data.shape[:-1] + (4,))
    r_rot_quat[..., 0] = torch.cos(r_rot_ang)
    r_rot_quat[..., 2] = torch.sin(r_rot_ang)
    r_pos = torch.zeros(data.shape[:-1] + (3,))
    r_pos[..., 1:, [0, 2]] = data[..., :-1, 1:3]
    r_pos = qrot(qinv(r_rot_quat), r_pos)
    r_pos = torch.cumsum(r_pos, dim=-2)
    r_pos[..., 1] = data[..., 3]
    return r_rot_quat, r_pos


def recover_from_ric(data: Tensor, joints_num: int = 22) -> Tensor:
    """mld/data/humanml/scripts/motion_process.py:415-431."""
    r_rot_quat, r_pos = recover_root_rot_pos(data)
    positions = data[..., 4:(joints_num - 1) * 3 + 4]
    positions = positions.reshape(positions.shape[:-1] + (-1, 3))
    positions = qrot(qinv(r_rot_quat[..., None, :]).expand(positions.shape[:-1] + (4,)), positions)
    positions = positions.clone()
    positions[..., 0] += r_pos[..., 0:1]
    positions[..., 2] += r_pos[..., 2:3]
    return torch.cat([r_pos.unsqueeze(-2), positions], dim=-2)


def feats2joints(feats: Tensor, mean: Tensor, std: Tensor, joints_num: int = 22) -> Tensor:
    """HumanML3DDataModule.feats2joints (mld/data/HumanML3D.py:41-45)."""
    return recover_from_ric(feats * std + mean, joints_num)


def remove_padding(tensors, lengths):
    """mld/utils/temos_utils.py:24-28."""
    return [t[:n] for t, n in zip(tensors, lengths)]


# --------------------------------------------------------------------------- MLD.forward
def mld_forward(dsd: SD, dcfg: DenoiserCfg, vsd: SD, vcfg: VaeCfg, scheduler,
                num_inference_steps: int, text_emb: Tensor, init_noise: Tensor,
                lengths: Sequence[int], mean: Tensor, std: Tensor,
                guidance_scale: float = 7.5, eta: float = 0.0):
    """MLD.forward after the text encoder (mld.py:231-265): _diffusion_reverse ->
    vae.decode -> feats2joints -> remove_padding.  Returns (joints list, feats, z)."""
    z = diffusion_reverse(dsd, dcfg, scheduler, num_inference_steps, text_emb, init_noise,
                          lengths, guidance_scale, eta)
    feats = vae_decode(vsd, vcfg, z, lengths)
    joints = feats2joints(feats, mean, std)
    return remove_padding(joints, lengths), feats, z
