"""ctypes binding of ``libmldb200.so`` (C ABI in ``include/mldb.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
Python/CPU fallback: a missing library is an ImportError with the build command, and every
non-zero status from the C side becomes a ``RuntimeError`` carrying ``mldb_last_error()``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmldb200.so")

MLDB_ABI_VERSION = 3
COND_TEXT, COND_ACTION = 0, 1
ARCH_TRANS_ENC, ARCH_TRANS_DEC = 0, 1
VAE_NONE, VAE_MLD, VAE_ACTOR = 0, 1, 2
SCHED_DDIM, SCHED_DDPM = 0, 1
DTYPE_F32 = 0
# mldb_kernel_stats indices (MLDB_KSTAT_* in include/mldb.h)
KSTAT_NAMES = ("gemm_tc", "gemm_ln_tc", "ffn_tc", "attn_tc", "attn_mma", "attn_simt", "gemm_simt", "ln_simt",
               "ln_unfused", "misc")


class MldbConfig(C.Structure):
    """``mldb_config`` (include/mldb.h)."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("cond_kind", C.c_int32), ("arch", C.c_int32), ("latent_dim", C.c_int32),
        ("n_lat", C.c_int32), ("num_heads", C.c_int32), ("ff_size", C.c_int32),
        ("num_layers", C.c_int32), ("text_dim", C.c_int32), ("nclasses", C.c_int32),
        ("nfeats", C.c_int32), ("diffusion_only", C.c_int32), ("flip_sin_to_cos", C.c_int32),
        ("freq_shift", C.c_float), ("guidance_scale", C.c_float),
        ("vae_kind", C.c_int32), ("vae_layers", C.c_int32), ("vae_heads", C.c_int32),
        ("vae_ff", C.c_int32), ("vae_nfeats", C.c_int32),
        ("sched_kind", C.c_int32), ("num_train_timesteps", C.c_int32),
        ("beta_start", C.c_double), ("beta_end", C.c_double), ("steps_offset", C.c_int32),
        ("set_alpha_to_one", C.c_int32), ("eta", C.c_float), ("njoints", C.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/mldb.h declares
_P = C.c_void_p
_SIGNATURES = {
    "mldb_default_config": (None, [C.POINTER(MldbConfig)]),
    "mldb_create": (C.c_int, [C.POINTER(MldbConfig), C.c_int, C.POINTER(_P)]),
    "mldb_destroy": (None, [_P]),
    "mldb_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int32, C.c_int32]),
    "mldb_finalize_weights": (C.c_int, [_P, _P]),
    "mldb_set_mean_std": (C.c_int, [_P, _P, _P, C.c_int32]),
    "mldb_scheduler_table": (C.c_int, [C.POINTER(MldbConfig), _P]),
    "mldb_scheduler_timesteps": (C.c_int, [C.POINTER(MldbConfig), C.c_int32, _P]),
    "mldb_scheduler_set_timesteps": (C.c_int, [_P, C.c_int32, _P]),
    "mldb_scheduler_step": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int64, _P, _P]),
    "mldb_denoise": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_diffusion_reverse": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_vae_decode": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "mldb_vae_encode": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "mldb_feats2joints": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "mldb_sample": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "mldb_sample_host": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_profile_op": (C.c_int, [_P, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "mldb_profile_steps": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mldb_debug_timeline": (C.c_int, [C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "mldb_debug_gemm": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_debug_ffn": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_debug_attention": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, _P, _P]),
    "mldb_comm_unique_id": (C.c_int, [_P]),
    "mldb_comm_init": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "mldb_comm_attach": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "mldb_comm_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mldb_allgather": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "mldb_sample_gather": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mldb_gather_wait": (C.c_int, [_P, _P]),
    "mldb_kernel_stats": (C.c_int, [_P, _P, C.c_int32]),
    "mldb_reset_kernel_stats": (C.c_int, [_P]),
    "mldb_last_error": (C.c_char_p, []),
    "mldb_abi_version": (C.c_int, []),
    "mldb_launch_count": (C.c_int64, [_P]),
    "mldb_set_option": (C.c_int, [_P, C.c_char_p, C.c_char_p]),
}

_lib = None


def lib():
    """Load libmldb200.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). mld_b200 has no CPU or PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)      # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if l.mldb_abi_version() != MLDB_ABI_VERSION:
            raise ImportError("libmldb200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(status: int, what: str = "mldb"):
    if status != 0:
        msg = lib().mldb_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def default_config() -> MldbConfig:
    cfg = MldbConfig()
    lib().mldb_default_config(C.byref(cfg))
    return cfg
