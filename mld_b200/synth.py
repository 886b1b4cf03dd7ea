"""Seeded synthetic weights and inputs with the reference's state-dict key names and shapes.

There are no checkpoints, CLIP weights or datasets offline, so tests, ``bench.py`` and
``smoke()`` run on random-init weights of the reference architecture.  Keys/shapes follow
the reference modules exactly (``MldDenoiser`` mld_denoiser.py:40-131, ``MldVae``
mld_vae.py:49-114, ``ActorVae`` actor_vae.py:26-51); ``oracle/make_golden.py`` loads these
dicts into the reference's own modules with ``strict=True``, which pins the key contract.
Biases and LayerNorm affines are randomised (the reference initialises them to 0/1) so that
parity tests exercise every term.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

Tensor = torch.Tensor


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def xavier(self, out_f: int, in_f: int) -> Tensor:
        a = math.sqrt(6.0 / (in_f + out_f))
        return (torch.rand(out_f, in_f, generator=self.g) * 2 - 1) * a

    def normal(self, *shape, std=1.0) -> Tensor:
        return torch.randn(*shape, generator=self.g) * std

    def uniform(self, *shape) -> Tensor:
        return torch.rand(*shape, generator=self.g)


def _attn(sd: Dict[str, Tensor], g: _Gen, p: str, d: int):
    sd[p + "in_proj_weight"] = g.xavier(3 * d, d)
    sd[p + "in_proj_bias"] = g.normal(3 * d, std=0.02)
    sd[p + "out_proj.weight"] = g.xavier(d, d)
    sd[p + "out_proj.bias"] = g.normal(d, std=0.02)


def _ln(sd, g: _Gen, p: str, d: int):
    sd[p + "weight"] = 1.0 + g.normal(d, std=0.1)
    sd[p + "bias"] = g.normal(d, std=0.05)


def _ffn(sd, g: _Gen, p: str, d: int, ff: int):
    sd[p + "linear1.weight"] = g.xavier(ff, d)
    sd[p + "linear1.bias"] = g.normal(ff, std=0.02)
    sd[p + "linear2.weight"] = g.xavier(d, ff)
    sd[p + "linear2.bias"] = g.normal(d, std=0.02)


def _enc_layer(sd, g, p, d, ff):
    _attn(sd, g, p + "self_attn.", d)
    _ffn(sd, g, p, d, ff)
    _ln(sd, g, p + "norm1.", d)
    _ln(sd, g, p + "norm2.", d)


def _dec_layer(sd, g, p, d, ff):
    _attn(sd, g, p + "self_attn.", d)
    _attn(sd, g, p + "multihead_attn.", d)
    _ffn(sd, g, p, d, ff)
    _ln(sd, g, p + "norm1.", d)
    _ln(sd, g, p + "norm2.", d)
    _ln(sd, g, p + "norm3.", d)


def _skip_stack(sd, g, p, d, ff, num_layers, layer_fn):
    nb = (num_layers - 1) // 2
    _ln(sd, g, p + "norm.", d)
    for i in range(nb):
        layer_fn(sd, g, f"{p}input_blocks.{i}.", d, ff)
    layer_fn(sd, g, f"{p}middle_block.", d, ff)
    for i in range(nb):
        layer_fn(sd, g, f"{p}output_blocks.{i}.", d, ff)
    for i in range(nb):
        sd[f"{p}linear_blocks.{i}.weight"] = g.xavier(d, 2 * d)
        sd[f"{p}linear_blocks.{i}.bias"] = g.normal(d, std=0.02)


def denoiser_state_dict(seed: int = 1234, condition: str = "text", arch: str = "trans_enc",
                        d: int = 256, ff: int = 1024, num_layers: int = 9,
                        text_dim: int = 768, nclasses: int = 12, nfeats: int = 263,
                        diffusion_only: bool = False) -> Dict[str, Tensor]:
    """State dict of ``MldDenoiser`` (text / action; skip trans_enc / no-VAE trans_dec)."""
    g, sd = _Gen(seed), {}
    if diffusion_only:
        sd["pose_embd.weight"] = g.xavier(d, nfeats)
        sd["pose_embd.bias"] = g.normal(d, std=0.02)
        sd["pose_proj.weight"] = g.xavier(nfeats, d)
        sd["pose_proj.bias"] = g.normal(nfeats, std=0.02)
    tdim = text_dim if condition == "text" else d
    sd["time_embedding.linear_1.weight"] = g.xavier(d, tdim)
    sd["time_embedding.linear_1.bias"] = g.normal(d, std=0.02)
    sd["time_embedding.linear_2.weight"] = g.xavier(d, d)
    sd["time_embedding.linear_2.bias"] = g.normal(d, std=0.02)
    if condition == "text":
        sd["emb_proj.1.weight"] = g.xavier(d, text_dim)
        sd["emb_proj.1.bias"] = g.normal(d, std=0.02)
    else:
        sd["emb_proj.action_embedding"] = g.xavier(nclasses, d)
    sd["query_pos.pe"] = g.uniform(500, 1, d)
    sd["mem_pos.pe"] = g.uniform(500, 1, d)
    if arch == "trans_enc":
        _skip_stack(sd, g, "encoder.", d, ff, num_layers, _enc_layer)
    else:
        for i in range(num_layers):
            _dec_layer(sd, g, f"decoder.layers.{i}.", d, ff)
        _ln(sd, g, "decoder.norm.", d)
    return sd


def mld_vae_state_dict(seed: int = 4321, nfeats: int = 263, d: int = 256, ff: int = 1024,
                       num_layers: int = 9, n_lat: int = 1) -> Dict[str, Tensor]:
    """State dict of ``MldVae`` (arch encoder_decoder, learned PE, MLP_DIST False)."""
    g, sd = _Gen(seed), {}
    sd["global_motion_token"] = g.normal(2 * n_lat, d)
    sd["query_pos_encoder.pe"] = g.uniform(500, 1, d)
    sd["query_pos_decoder.pe"] = g.uniform(500, 1, d)
    _skip_stack(sd, g, "encoder.", d, ff, num_layers, _enc_layer)
    _skip_stack(sd, g, "decoder.", d, ff, num_layers, _dec_layer)
    sd["skel_embedding.weight"] = g.xavier(d, nfeats)
    sd["skel_embedding.bias"] = g.normal(d, std=0.02)
    sd["final_layer.weight"] = g.xavier(nfeats, d)
    sd["final_layer.bias"] = g.normal(nfeats, std=0.02)
    return sd


def sine_pe_table(n: int, d: int) -> Tensor:
    """The ``PositionalEncoding`` buffer (position_encoding_layer.py:14-20), rows [n, d]."""
    pe = torch.zeros(n, d)
    position = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def actor_vae_state_dict(seed: int = 777, nfeats: int = 150, d: int = 256, ff: int = 1024,
                         num_layers: int = 6, with_encoder: bool = True) -> Dict[str, Tensor]:
    """State dict of ``ActorVae`` (torch nn.TransformerEncoder/Decoder stacks, sine PE)."""
    g, sd = _Gen(seed), {}
    pe = sine_pe_table(5000, d).unsqueeze(1)
    if with_encoder:
        sd["encoder.mu_token"] = g.normal(d)
        sd["encoder.logvar_token"] = g.normal(d)
        sd["encoder.skel_embedding.weight"] = g.xavier(d, nfeats)
        sd["encoder.skel_embedding.bias"] = g.normal(d, std=0.02)
        sd["encoder.sequence_pos_encoding.pe"] = pe.clone()
        for i in range(num_layers):
            _enc_layer(sd, g, f"encoder.seqTransEncoder.layers.{i}.", d, ff)
    sd["decoder.sequence_pos_encoding.pe"] = pe.clone()
    for i in range(num_layers):
        _dec_layer(sd, g, f"decoder.seqTransDecoder.layers.{i}.", d, ff)
    sd["decoder.final_layer.weight"] = g.xavier(nfeats, d)
    sd["decoder.final_layer.bias"] = g.normal(nfeats, std=0.02)
    return sd


def text_context(B: int, S: int, seed: int = 1, text_dim: int = 768) -> Tensor:
    """Synthetic CLIP output ``[2B, S, 768]`` for CFG: the uncond half first
    (mld.py:225-230), all uncond rows equal (every "" prompt encodes identically)."""
    g = torch.Generator().manual_seed(seed)
    uncond = torch.randn(1, S, text_dim, generator=g).expand(B, S, text_dim)
    cond = torch.randn(B, S, text_dim, generator=g)
    return torch.cat([uncond, cond], 0).contiguous()


def init_noise(B: int, n_lat: int = 1, d: int = 256, seed: int = 2) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, n_lat, d, generator=g)


def ragged_lengths(B: int, lo: int = 40, hi: int = 196, seed: int = 3) -> List[int]:
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(lo, hi + 1, (B,), generator=g) // 4 * 4).tolist()


def mean_std(nfeats: int = 263, seed: int = 5):
    """Synthetic dataset statistics (``Mean.npy`` / ``Std.npy`` are absent offline) with
    HumanML3D-like magnitudes: per-frame root rotation / translation velocities of a few
    hundredths (rad, m), root height ~0.9 m, joint offsets of decimetres.  feats2joints
    integrates the velocities over up to 196 frames, so the magnitudes matter for how the
    1e-3 joint-position gate conditions the path."""
    g = torch.Generator().manual_seed(seed)
    mean = torch.randn(nfeats, generator=g) * 0.05
    std = 0.2 + 0.8 * torch.rand(nfeats, generator=g)
    mean[0], std[0] = 0.0, 0.03                      # root angular velocity (rad/frame)
    mean[1:3] = torch.tensor([0.0, 0.02])
    std[1:3] = 0.03                                  # root linear velocity xz (m/frame)
    mean[3], std[3] = 0.9, 0.1                       # root height
    n_ric = min(nfeats, 67) - 4
    if n_ric > 0:
        mean[4:4 + n_ric] = torch.randn(n_ric, generator=g) * 0.3
        std[4:4 + n_ric] = 0.1 + 0.2 * torch.rand(n_ric, generator=g)
    return mean, std
