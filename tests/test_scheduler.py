"""Scheduler known-answer tests.  diffusers is absent from the reference tree (third-party,
unpinned), so the restatement is anchored on (i) the published DDIM/DDPM update rules,
(ii) the reference's configuration (configs/modules/scheduler.yaml) and call sites
(mld.py:310-320,345), (iii) hand-derived known answers.  The C library's host-side tables are
checked bit-exact against torch arithmetic (no GPU needed)."""
import ctypes as C

import numpy as np
import torch

from mld_b200 import _lib
from oracle import mld_oracle as O


def test_ddim_timesteps_known_answer():
    s = O.DDIMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.dtype == torch.int64
    assert s.timesteps.tolist() == list(range(981, 0, -20))            # [981, 961, ..., 21, 1]
    s.set_timesteps(1000)
    assert s.timesteps[0].item() == 1000 - 1 + 1 and s.timesteps[-1].item() == 1
    d = O.DDPMScheduler()
    d.set_timesteps(1000)
    assert d.timesteps.tolist() == list(range(999, -1, -1))


def test_alphas_cumprod_known_values():
    s = O.DDIMScheduler()
    ac = s.alphas_cumprod
    assert ac.dtype == torch.float32 and ac.shape == (1000,)
    # scaled_linear: beta_0 = 0.00085, beta_999 = 0.012
    assert abs(float(1 - ac[0]) - 0.00085) < 1e-7
    assert abs(float(ac[999] / ac[998]) - (1 - 0.012)) < 1e-6
    ref = torch.cumprod(1 - torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2, 0)
    assert float((ac.double() - ref).abs().max()) < 1e-6
    assert float(s.final_alpha_cumprod) == float(ac[0])               # set_alpha_to_one: false


def test_ddim_step_is_exact_inverse_of_forward_noising():
    """x_t = sqrt(a_t) x0 + sqrt(1-a_t) eps; one eta=0 step with the true eps must land on
    sqrt(a_prev) x0 + sqrt(1-a_prev) eps (the defining property of DDIM)."""
    s = O.DDIMScheduler()
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x0, eps = torch.randn(4, 1, 256, generator=g), torch.randn(4, 1, 256, generator=g)
    for t in (981, 501, 21):
        a_t, a_p = s.alphas_cumprod[t], s.alphas_cumprod[t - 20]
        x_t = a_t.sqrt() * x0 + (1 - a_t).sqrt() * eps
        want = a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps
        assert torch.allclose(s.step(eps, t, x_t), want, atol=2e-5)
    # last step (t=1): prev_t < 0 -> final_alpha_cumprod = alphas_cumprod[0]
    a_t, a_p = s.alphas_cumprod[1], s.alphas_cumprod[0]
    x_t = a_t.sqrt() * x0 + (1 - a_t).sqrt() * eps
    assert torch.allclose(s.step(eps, 1, x_t), a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps, atol=2e-5)


def test_ddpm_step_posterior_mean_and_variance():
    d = O.DDPMScheduler()
    d.set_timesteps(1000)
    g = torch.Generator().manual_seed(1)
    x0, eps, nz = (torch.randn(2, 8, 263, generator=g) for _ in range(3))
    t = 500
    ac = d.alphas_cumprod.double()
    a_t, a_p = ac[t], ac[t - 1]
    beta = 1 - a_t / a_p
    x_t = a_t.sqrt() * x0.double() + (1 - a_t).sqrt() * eps.double()
    mean = (a_p.sqrt() * beta / (1 - a_t)) * x0.double() + ((a_t / a_p).sqrt() * (1 - a_p) / (1 - a_t)) * x_t
    var = (1 - a_p) / (1 - a_t) * beta
    want = mean + var.sqrt() * nz.double()
    got = d.step(eps, t, x_t.float(), noise=nz)
    assert torch.allclose(got.double(), want, atol=1e-4)
    # t == 0: no noise is added
    assert torch.equal(d.step(eps, 0, x_t.float(), noise=nz), d.step(eps, 0, x_t.float(), noise=None))


def test_library_tables_bit_exact(built_lib):
    cfg = _lib.default_config()
    ac = torch.empty(1000, dtype=torch.float32)
    assert built_lib.mldb_scheduler_table(C.byref(cfg), C.c_void_p(ac.data_ptr())) == 0
    assert torch.equal(ac, O.DDIMScheduler().alphas_cumprod)          # bit-exact with torch fp32
    for n in (50, 1000, 200, 25):
        ts = torch.empty(n, dtype=torch.int64)
        assert built_lib.mldb_scheduler_timesteps(C.byref(cfg), n, C.c_void_p(ts.data_ptr())) == 0
        s = O.DDIMScheduler()
        s.set_timesteps(n)
        assert torch.equal(ts, s.timesteps)                           # int64 equality
    cfg.sched_kind = _lib.SCHED_DDPM
    ts = torch.empty(1000, dtype=torch.int64)
    built_lib.mldb_scheduler_timesteps(C.byref(cfg), 1000, C.c_void_p(ts.data_ptr()))
    assert ts.tolist() == list(range(999, -1, -1))
