"""Thin torch-side wrapper over the C ABI: one ``Engine`` == one ``mldb_handle`` on one GPU.

PyTorch is plumbing only here: it owns the device tensors and the current CUDA stream; every
FLOP of the sampling path runs inside ``libmldb200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from ._lib import MldbConfig, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """Owns an ``mldb_handle``.  ``cfg`` is an :class:`MldbConfig` (see ``make_config``)."""

    def __init__(self, cfg: MldbConfig, device: int | torch.device = 0):
        self._h = None
        self.lib = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("mld_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.device = dev
        self.cfg = cfg
        h = C.c_void_p()
        check(self.lib.mldb_create(C.byref(cfg), dev.index or 0, C.byref(h)), "mldb_create")
        self._h = h
        self.timesteps: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if self._h is not None:
                self.lib.mldb_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        """Feed every tensor of a reference state dict (keys get ``prefix``, e.g. 'denoiser.')."""
        for k, v in sd.items():
            t = v.detach().to(dtype=torch.float32).contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(self.lib.mldb_load_tensor(self._h, (prefix + k).encode(), _ptr(t), shape, t.dim(),
                                            _lib.DTYPE_F32), f"mldb_load_tensor({prefix + k})")

    def finalize(self):
        check(self.lib.mldb_finalize_weights(self._h, None), "mldb_finalize_weights")

    def set_mean_std(self, mean: torch.Tensor, std: torch.Tensor):
        m = mean.detach().float().contiguous().cpu()
        s = std.detach().float().contiguous().cpu()
        check(self.lib.mldb_set_mean_std(self._h, _ptr(m), _ptr(s), m.numel()), "mldb_set_mean_std")

    def set_option(self, name: str, value: str):
        check(self.lib.mldb_set_option(self._h, name.encode(), value.encode()), "mldb_set_option")

    def profile_op(self, op: str, B: int, S_ctx: int, iters: int = 20) -> float:
        """Average ms of one operator of denoiser layer 0 in isolation (bench.py roofline leg)."""
        ms = C.c_float()
        check(self.lib.mldb_profile_op(self._h, op.encode(), B, S_ctx, iters, C.byref(ms)), "mldb_profile_op")
        return float(ms.value)

    def debug_gemm(self, A, W, bias=None, gamma=None, beta=None, R=None, K1=0, act=0, use_tc=True, split_out=False):
        """Kernel unit-test hook (mldb_debug_gemm): A [M,K] (device), W [N,K] / bias / gamma / beta (host)."""
        A = _f32c(A, self.device)
        Wc = W.detach().float().contiguous().cpu()
        host = [None if t is None else t.detach().float().contiguous().cpu() for t in (bias, gamma, beta)]
        Rd = None if R is None else _f32c(R, self.device)
        M, K = A.shape
        N = Wc.shape[0]
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_debug_gemm(self._h, _ptr(A), _ptr(Wc), _ptr(host[0]), _ptr(host[1]), _ptr(host[2]),
                                       _ptr(Rd), M, N, K, K1, act, int(use_tc), int(split_out), _ptr(out), self._stream()),
              "mldb_debug_gemm")
        return out

    def debug_ffn(self, X, W1, b1, W2, b2, gamma, beta, mode=2):
        """Kernel unit-test hook (mldb_debug_ffn): LayerNorm(X + W2 gelu(W1 X + b1) + b2); mode 0 CUDA-core,
        1 tcgen05 GEMMs (two launches), 2 fused tcgen05 FFN kernel.  X [M,d] (device), the rest host."""
        X = _f32c(X, self.device)
        host = [t.detach().float().contiguous().cpu() for t in (W1, b1, W2, b2, gamma, beta)]
        M, d = X.shape
        ff = host[0].shape[0]
        out = torch.empty((M, d), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_debug_ffn(self._h, _ptr(X), *[_ptr(t) for t in host], M, d, ff, int(mode), _ptr(out),
                                      self._stream()), "mldb_debug_ffn")
        return out

    def debug_attention(self, qkv, nseq, L, heads, lengths=None, mode=1):
        """Kernel unit-test hook (mldb_debug_attention): qkv [nseq*L, 3*heads*hd] (device), mode 0 CUDA-core,
        1 mma.sync (product), 2 tcgen05 (experimental).  Returns [nseq*L, heads*hd]."""
        qkv = _f32c(qkv, self.device)
        d = qkv.shape[1] // 3
        ln = None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=self.device).contiguous()
        out = torch.empty((nseq * L, d), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_debug_attention(self._h, _ptr(qkv), _ptr(ln), nseq, L, heads, d // heads, int(mode), _ptr(out),
                                            self._stream()), "mldb_debug_attention")
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.mldb_launch_count(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ scheduler
    def set_timesteps(self, n: int) -> torch.Tensor:
        ts = torch.empty(n, dtype=torch.int64)
        check(self.lib.mldb_scheduler_set_timesteps(self._h, n, _ptr(ts)), "mldb_scheduler_set_timesteps")
        self.timesteps = ts
        return ts

    def scheduler_step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor,
                       noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        mo, sa = _f32c(model_output, self.device), _f32c(sample, self.device)
        nz = None if noise is None else _f32c(noise, self.device)
        out = torch.empty_like(sa)
        check(self.lib.mldb_scheduler_step(self._h, _ptr(mo), int(timestep), _ptr(sa), _ptr(nz),
                                           sa.numel(), _ptr(out), self._stream()), "mldb_scheduler_step")
        return out

    # ------------------------------------------------------------------ denoiser
    def _cond(self, cond: torch.Tensor) -> torch.Tensor:
        if self.cfg.cond_kind == _lib.COND_TEXT:
            return _f32c(cond, self.device)
        return cond.to(device=self.device, dtype=torch.int64).contiguous()

    def _lengths(self, lengths) -> Optional[torch.Tensor]:
        if lengths is None:
            return None
        if isinstance(lengths, torch.Tensor):
            return lengths.to(device=self.device, dtype=torch.int32).contiguous()
        return torch.tensor(list(lengths), dtype=torch.int32, device=self.device)

    def denoise(self, sample: torch.Tensor, timestep: int, cond: torch.Tensor,
                lengths=None) -> torch.Tensor:
        x, c = _f32c(sample, self.device), self._cond(cond)
        ln = self._lengths(lengths)
        Bx = x.shape[0]
        S = c.shape[1] if c.dim() == 3 else 1
        T = x.shape[1] if self.cfg.diffusion_only else 0
        out = torch.empty_like(x)
        check(self.lib.mldb_denoise(self._h, _ptr(x), int(timestep), _ptr(c), _ptr(ln), Bx, S, T,
                                    _ptr(out), self._stream()), "mldb_denoise")
        return out

    def diffusion_reverse(self, cond: torch.Tensor, init_noise: torch.Tensor, lengths=None,
                          step_noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        c, z0 = self._cond(cond), _f32c(init_noise, self.device)
        ln = self._lengths(lengths)
        B = z0.shape[0]
        S = c.shape[1] if c.dim() == 3 else 1
        T = z0.shape[1] if self.cfg.diffusion_only else 0
        out = torch.empty((z0.shape[1], B, z0.shape[2]), dtype=torch.float32, device=self.device)
        sn = None if step_noise is None else _f32c(step_noise, self.device)
        check(self.lib.mldb_diffusion_reverse(self._h, _ptr(c), _ptr(z0), _ptr(sn), _ptr(ln), B, S, T,
                                              _ptr(out), self._stream()), "mldb_diffusion_reverse")
        return out

    # ------------------------------------------------------------------ VAE / joints
    def vae_decode(self, z: torch.Tensor, lengths) -> torch.Tensor:
        zz = _f32c(z, self.device)
        ln = self._lengths(lengths)
        B = zz.shape[1]
        T = int(max(lengths)) if not isinstance(lengths, torch.Tensor) else int(lengths.max())
        out = torch.empty((B, T, self.cfg.vae_nfeats), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_vae_decode(self._h, _ptr(zz), _ptr(ln), B, T, _ptr(out), self._stream()),
              "mldb_vae_decode")
        return out

    def vae_encode(self, feats: torch.Tensor, lengths):
        f = _f32c(feats, self.device)
        ln = self._lengths(lengths)
        B, T = f.shape[0], f.shape[1]
        shape = (self.cfg.n_lat, B, self.cfg.latent_dim)
        mu = torch.empty(shape, dtype=torch.float32, device=self.device)
        logvar = torch.empty(shape, dtype=torch.float32, device=self.device)
        check(self.lib.mldb_vae_encode(self._h, _ptr(f), _ptr(ln), B, T, _ptr(mu), _ptr(logvar),
                                       self._stream()), "mldb_vae_encode")
        return mu, logvar

    def feats2joints(self, feats: torch.Tensor) -> torch.Tensor:
        f = _f32c(feats, self.device)
        B, T = f.shape[0], f.shape[1]
        out = torch.empty((B, T, self.cfg.njoints, 3), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_feats2joints(self._h, _ptr(f), B, T, _ptr(out), self._stream()),
              "mldb_feats2joints")
        return out

    # ------------------------------------------------------------------ fused sample
    def sample(self, cond: torch.Tensor, init_noise: torch.Tensor, lengths, want=("joints",)):
        """reverse diffusion -> decode -> feats2joints on device tensors.  Returns a dict with the
        requested subset of {"latents" [n_lat,B,d], "feats" [B,T,F], "joints" [B,T,J,3]}."""
        c, z0 = self._cond(cond), _f32c(init_noise, self.device)
        ln = self._lengths(lengths)
        B = z0.shape[0]
        S = c.shape[1] if c.dim() == 3 else 1
        T = int(max(lengths)) if not isinstance(lengths, torch.Tensor) else int(lengths.max())
        cfg = self.cfg
        out = {}
        lat = fe = jo = None
        if "latents" in want:
            lat = out["latents"] = torch.empty((cfg.n_lat, B, cfg.latent_dim), dtype=torch.float32, device=self.device)
        if "feats" in want:
            fe = out["feats"] = torch.empty((B, T, cfg.vae_nfeats), dtype=torch.float32, device=self.device)
        if "joints" in want:
            jo = out["joints"] = torch.empty((B, T, cfg.njoints, 3), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_sample(self._h, _ptr(c), _ptr(z0), _ptr(ln), B, S, T, _ptr(lat), _ptr(fe),
                                   _ptr(jo), self._stream()), "mldb_sample")
        return out

    def sample_host(self, cond_cpu: torch.Tensor, noise_cpu: torch.Tensor, lengths_cpu: torch.Tensor,
                    joints_cpu: torch.Tensor, T: int):
        """End-to-end through HOST buffers (pinned recommended); asynchronous on the current
        stream - synchronise before reading ``joints_cpu``."""
        B = noise_cpu.shape[0]
        S = cond_cpu.shape[1] if cond_cpu.dim() == 3 else 1
        assert lengths_cpu.dtype == torch.int32 and joints_cpu.dtype == torch.float32
        check(self.lib.mldb_sample_host(self._h, _ptr(cond_cpu), _ptr(noise_cpu), _ptr(lengths_cpu), B, S, T,
                                        _ptr(joints_cpu), self._stream()), "mldb_sample_host")
        return joints_cpu


def make_config(*, condition: str = "text", arch: str = "trans_enc", latent_dim: Sequence[int] = (1, 256),
                ff_size: int = 1024, num_layers: int = 9, num_heads: int = 4, text_encoded_dim: int = 768,
                nclasses: int = 12, nfeats: int = 263, diffusion_only: bool = False,
                flip_sin_to_cos: bool = True, freq_shift: float = 0.0, guidance_scale: float = 7.5,
                vae: str = "mld", vae_layers: Optional[int] = None, vae_heads: int = 4, vae_ff: int = 1024,
                vae_nfeats: Optional[int] = None, scheduler: str = "ddim", num_train_timesteps: int = 1000,
                beta_start: float = 0.00085, beta_end: float = 0.012, steps_offset: int = 1,
                set_alpha_to_one: bool = False, eta: float = 0.0, njoints: int = 22) -> MldbConfig:
    """Build an ``mldb_config`` from the reference's ctor kwargs / yaml params
    (configs/modules/{denoiser,motion_vae,scheduler}.yaml)."""
    c = _lib.default_config()
    c.cond_kind = {"text": _lib.COND_TEXT, "action": _lib.COND_ACTION}[condition]
    c.arch = {"trans_enc": _lib.ARCH_TRANS_ENC, "trans_dec": _lib.ARCH_TRANS_DEC}[arch]
    c.n_lat, c.latent_dim = int(latent_dim[0]), int(latent_dim[-1])
    c.ff_size, c.num_layers, c.num_heads = ff_size, num_layers, num_heads
    c.text_dim, c.nclasses, c.nfeats = text_encoded_dim, nclasses, nfeats
    c.diffusion_only = int(diffusion_only)
    c.flip_sin_to_cos, c.freq_shift, c.guidance_scale = int(flip_sin_to_cos), freq_shift, guidance_scale
    c.vae_kind = {"none": _lib.VAE_NONE, "no": _lib.VAE_NONE, "mld": _lib.VAE_MLD, "actor": _lib.VAE_ACTOR}[vae]
    c.vae_layers = vae_layers if vae_layers is not None else (6 if vae == "actor" else 9)
    c.vae_heads, c.vae_ff = vae_heads, vae_ff
    c.vae_nfeats = vae_nfeats if vae_nfeats is not None else nfeats
    c.sched_kind = {"ddim": _lib.SCHED_DDIM, "ddpm": _lib.SCHED_DDPM}[scheduler]
    c.num_train_timesteps, c.beta_start, c.beta_end = num_train_timesteps, beta_start, beta_end
    c.steps_offset, c.set_alpha_to_one, c.eta, c.njoints = steps_offset, int(set_alpha_to_one), eta, njoints
    return c
