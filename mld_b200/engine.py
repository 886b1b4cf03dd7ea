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

    def profile_steps(self, cond: torch.Tensor, init_noise: torch.Tensor):
        """Device time (ms) of every scheduler step of the reverse loop, launched eagerly (mldb_profile_steps)."""
        c, z0 = self._cond(cond), _f32c(init_noise, self.device)
        B = z0.shape[0]
        self._check_latent(z0, B, "init_noise")
        self._check_cond(c, 2 * B if self.cfg_on else B)
        S = c.shape[1] if c.dim() == 3 else 1
        n = 0 if self.timesteps is None else len(self.timesteps)
        ms = (C.c_float * n)()
        check(self.lib.mldb_profile_steps(self._h, _ptr(c), _ptr(z0), B, S, ms), "mldb_profile_steps")
        return [float(v) for v in ms]

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

    def debug_attention(self, q, nseq, Lq, heads, lengths=None, mode=2, kv=None, Lk=None, kv_prefix=0):
        """Kernel unit-test hook (mldb_debug_attention).  ``kv is None``: ``q`` is a packed qkv
        [nseq*Lq, 3*heads*hd] tensor (self-attention, the layout the stacks use); else ``q`` [nseq*Lq, d] and
        ``kv`` [nseq*Lk, 2*d] (cross-attention).  mode 0 CUDA-core, 1 mma.sync, 2 tcgen05 (product).
        Returns [nseq*Lq, heads*hd]."""
        q = _f32c(q, self.device)
        kvd = None if kv is None else _f32c(kv, self.device)
        d = q.shape[1] // 3 if kv is None else q.shape[1]
        Lk = Lq if Lk is None else Lk
        if q.shape[0] != nseq * Lq or (kvd is not None and tuple(kvd.shape) != (nseq * Lk, 2 * d)):
            raise ValueError("debug_attention: shape mismatch")
        ln = None if lengths is None else torch.as_tensor(lengths, dtype=torch.int32, device=self.device).contiguous()
        out = torch.empty((nseq * Lq, d), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_debug_attention(self._h, _ptr(q), _ptr(kvd), _ptr(ln), int(kv_prefix), nseq, Lq, Lk, heads,
                                            d // heads, int(mode), _ptr(out), self._stream()), "mldb_debug_attention")
        return out

    def kernel_stats(self, reset: bool = False) -> Dict[str, int]:
        """Which kernel each operator was enqueued on since the last reset (mldb_kernel_stats)."""
        arr = (C.c_int64 * len(_lib.KSTAT_NAMES))()
        check(self.lib.mldb_kernel_stats(self._h, arr, len(_lib.KSTAT_NAMES)), "mldb_kernel_stats")
        if reset:
            check(self.lib.mldb_reset_kernel_stats(self._h), "mldb_reset_kernel_stats")
        return {k: int(v) for k, v in zip(_lib.KSTAT_NAMES, arr)}

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

    # ------------------------------------------------------------------ argument checks
    # The C ABI takes raw pointers: a wrong shape would be a silent out-of-bounds access on the device.
    @property
    def cfg_on(self) -> bool:
        return self.cfg.guidance_scale > 1.0

    def _check_cond(self, c: torch.Tensor, Bx: int):
        if self.cfg.cond_kind == _lib.COND_TEXT:
            if c.dim() != 3 or c.shape[0] != Bx or c.shape[2] != self.cfg.text_dim or c.shape[1] < 1:
                raise ValueError(f"condition must be [{Bx}, S, {self.cfg.text_dim}] (uncond half first when guidance "
                                 f"is on), got {tuple(c.shape)}")
        elif c.numel() != Bx:
            raise ValueError(f"action condition must hold {Bx} class ids ([{Bx}, 1]), got {tuple(c.shape)}")

    def _check_latent(self, x: torch.Tensor, rows: int, what: str):
        cfg = self.cfg
        want = (rows, x.shape[1], cfg.nfeats) if cfg.diffusion_only else (rows, cfg.n_lat, cfg.latent_dim)
        if x.dim() != 3 or tuple(x.shape) != want:
            raise ValueError(f"{what} must be {list(want)}, got {tuple(x.shape)}")

    def _check_lengths(self, ln: Optional[torch.Tensor], B: int, T: Optional[int] = None):
        if ln is None:
            return
        if ln.numel() != B:
            raise ValueError(f"lengths must have {B} entries, got {ln.numel()}")
        if T is not None and not isinstance(T, bool):
            host = ln if ln.device.type == "cpu" else None
            if host is not None and (int(host.max()) > T or int(host.min()) < 1):
                raise ValueError(f"lengths must lie in [1, {T}]")

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
        self._check_latent(x, Bx, "sample")
        self._check_cond(c, Bx)
        self._check_lengths(ln, Bx)
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
        self._check_latent(z0, B, "init_noise")
        self._check_cond(c, 2 * B if self.cfg_on else B)
        self._check_lengths(ln, B)
        S = c.shape[1] if c.dim() == 3 else 1
        T = z0.shape[1] if self.cfg.diffusion_only else 0
        out = torch.empty((z0.shape[1], B, z0.shape[2]), dtype=torch.float32, device=self.device)
        sn = None if step_noise is None else _f32c(step_noise, self.device)
        if sn is not None:
            n_steps = 0 if self.timesteps is None else len(self.timesteps)
            if tuple(sn.shape) != (n_steps, *z0.shape):
                raise ValueError(f"step_noise must be [{n_steps}, {', '.join(map(str, z0.shape))}], got {tuple(sn.shape)}")
        check(self.lib.mldb_diffusion_reverse(self._h, _ptr(c), _ptr(z0), _ptr(sn), _ptr(ln), B, S, T,
                                              _ptr(out), self._stream()), "mldb_diffusion_reverse")
        return out

    # ------------------------------------------------------------------ VAE / joints
    def vae_decode(self, z: torch.Tensor, lengths) -> torch.Tensor:
        zz = _f32c(z, self.device)
        ln = self._lengths(lengths)
        if zz.dim() != 3 or zz.shape[0] != self.cfg.n_lat or zz.shape[2] != self.cfg.latent_dim:
            raise ValueError(f"z must be [{self.cfg.n_lat}, B, {self.cfg.latent_dim}], got {tuple(zz.shape)}")
        B = zz.shape[1]
        self._check_lengths(ln, B)
        T = int(max(lengths)) if not isinstance(lengths, torch.Tensor) else int(lengths.max())
        out = torch.empty((B, T, self.cfg.vae_nfeats), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_vae_decode(self._h, _ptr(zz), _ptr(ln), B, T, _ptr(out), self._stream()),
              "mldb_vae_decode")
        return out

    def vae_encode(self, feats: torch.Tensor, lengths):
        f = _f32c(feats, self.device)
        ln = self._lengths(lengths)
        if f.dim() != 3 or f.shape[2] != self.cfg.vae_nfeats:
            raise ValueError(f"feats must be [B, T, {self.cfg.vae_nfeats}], got {tuple(f.shape)}")
        B, T = f.shape[0], f.shape[1]
        self._check_lengths(ln, B)
        shape = (self.cfg.n_lat, B, self.cfg.latent_dim)
        mu = torch.empty(shape, dtype=torch.float32, device=self.device)
        logvar = torch.empty(shape, dtype=torch.float32, device=self.device)
        check(self.lib.mldb_vae_encode(self._h, _ptr(f), _ptr(ln), B, T, _ptr(mu), _ptr(logvar),
                                       self._stream()), "mldb_vae_encode")
        return mu, logvar

    def feats2joints(self, feats: torch.Tensor) -> torch.Tensor:
        f = _f32c(feats, self.device)
        F = self.cfg.vae_nfeats if self.cfg.vae_kind != _lib.VAE_NONE else self.cfg.nfeats
        if f.dim() != 3 or f.shape[2] != F:
            raise ValueError(f"feats must be [B, T, {F}], got {tuple(f.shape)}")
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
        self._check_latent(z0, B, "init_noise")
        self._check_cond(c, 2 * B if self.cfg_on else B)
        self._check_lengths(ln, B)
        S = c.shape[1] if c.dim() == 3 else 1
        T = int(max(lengths)) if not isinstance(lengths, torch.Tensor) else int(lengths.max())
        if T < 1 or int(min(lengths)) < 1:
            raise ValueError("lengths must be >= 1")
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

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    def comm_init(self, group=None):
        """Build the handle's NCCL communicator over the ranks of ``group`` (default: the world): rank 0 of
        the group draws the unique id inside the library, torch.distributed only ships its 128 bytes."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        uid = (C.c_ubyte * 128)()
        if rank == 0:
            check(self.lib.mldb_comm_unique_id(uid), "mldb_comm_unique_id")
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        buf = (C.c_ubyte * 128).from_buffer_copy(box[0])
        check(self.lib.mldb_comm_init(self._h, buf, world, rank), "mldb_comm_init")
        return world, rank

    def comm_info(self):
        n, r = C.c_int32(), C.c_int32()
        check(self.lib.mldb_comm_info(self._h, C.byref(n), C.byref(r)), "mldb_comm_info")
        return int(n.value), int(r.value)

    def allgather(self, local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ncclAllGather through the C ABI on the current stream: [n, ...] per rank -> [world * n, ...]."""
        world, _ = self.comm_info()
        x = _f32c(local, self.device)
        if out is None:
            out = torch.empty((world * x.shape[0], *x.shape[1:]), dtype=torch.float32, device=self.device)
        check(self.lib.mldb_allgather(self._h, _ptr(x), _ptr(out), x.numel(), self._stream()), "mldb_allgather")
        return out

    def sample_gather(self, cond: torch.Tensor, init_noise: torch.Tensor, lengths, T: Optional[int] = None,
                      out: Optional[torch.Tensor] = None, wait: bool = True) -> torch.Tensor:
        """This rank's shard through ``mldb_sample_gather``: joints of ALL ranks ``[world * B, T, J, 3]``.
        ``T`` must be the same on every rank (pad to the global max length).  ``wait=False`` leaves the gather
        running on the side stream (call :meth:`gather_wait` before reading ``out``; alternate two ``out``
        buffers between consecutive calls)."""
        c, z0 = self._cond(cond), _f32c(init_noise, self.device)
        ln = self._lengths(lengths)
        B = z0.shape[0]
        self._check_latent(z0, B, "init_noise")
        self._check_cond(c, 2 * B if self.cfg_on else B)
        self._check_lengths(ln, B)
        S = c.shape[1] if c.dim() == 3 else 1
        if T is None:
            T = int(max(lengths)) if not isinstance(lengths, torch.Tensor) else int(lengths.max())
        world, _ = self.comm_info()
        shape = (world * B, T, self.cfg.njoints, 3)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 {list(shape)} tensor")
        check(self.lib.mldb_sample_gather(self._h, _ptr(c), _ptr(z0), _ptr(ln), B, S, T, _ptr(out), self._stream()),
              "mldb_sample_gather")
        if wait:
            self.gather_wait()
        return out

    def gather_wait(self):
        check(self.lib.mldb_gather_wait(self._h, self._stream()), "mldb_gather_wait")

    def sample_host(self, cond_cpu: torch.Tensor, noise_cpu: torch.Tensor, lengths_cpu: torch.Tensor,
                    joints_cpu: torch.Tensor, T: int):
        """End-to-end through HOST buffers (pinned recommended); asynchronous on the current
        stream - synchronise before reading ``joints_cpu``."""
        B = noise_cpu.shape[0]
        S = cond_cpu.shape[1] if cond_cpu.dim() == 3 else 1
        for t, dt in ((cond_cpu, torch.float32 if self.cfg.cond_kind == _lib.COND_TEXT else torch.int64),
                      (noise_cpu, torch.float32), (lengths_cpu, torch.int32), (joints_cpu, torch.float32)):
            if t.device.type != "cpu" or t.dtype != dt or not t.is_contiguous():
                raise ValueError("sample_host takes contiguous HOST tensors (cond f32/int64, noise f32, lengths int32, joints f32)")
        self._check_latent(noise_cpu, B, "init_noise")
        self._check_cond(cond_cpu, 2 * B if self.cfg_on else B)
        self._check_lengths(lengths_cpu, B, T)
        world, _ = self.comm_info()
        if joints_cpu.numel() != world * B * T * self.cfg.njoints * 3:
            raise ValueError(f"joints buffer must hold [{world * B}, {T}, {self.cfg.njoints}, 3] floats")
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
