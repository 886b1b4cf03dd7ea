"""GPU kernel unit tests (pytest -m gpu): each GEMM epilogue of the tcgen05 kernel against a
float64 torch reference of the same op, and against the CUDA-core kernel, through the C ABI's
debug hook.  Tolerance 5e-6 relative-to-max: the split-fp16 3-product scheme keeps ~22 bits."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def eng(built_lib):
    from mld_b200.engine import Engine, make_config
    return Engine(make_config(num_layers=0, vae="none"), 0)


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


CASES = [
    # M,    N,    K,    K1,  act, ln
    (200, 768, 256, 0, 0, False),       # QKV projection, ragged M tail
    (1000, 1024, 256, 0, 1, False),     # FFN up-projection + exact GELU
    (333, 256, 1024, 0, 0, True),       # FFN down-projection + residual + LayerNorm (fused epilogue)
    (640, 256, 256, 0, 0, True),        # attention out-projection + residual + LayerNorm
    (515, 256, 512, 256, 0, False),     # skip connection: cat([x, skip]) @ W^T as two A sources
    (392, 263, 256, 0, 0, False),       # final layer: N not a tile multiple (TMA zero fill)
    (129, 512, 256, 0, 2, False),       # kv projection, ReLU epilogue variant
    (4096, 1024, 256, 0, 3, False),     # many tiles, SiLU
    (1300, 256, 1024, 0, 0, True),      # 11 m-tiles: cluster groups with an out-of-range m-tile, LN epilogue
    (1100, 263, 512, 256, 0, False),    # clusters + ragged N + two A sources
]


@pytest.mark.parametrize("M,N,K,K1,act,ln", CASES)
def test_tc_gemm_epilogues(eng, M, N, K, K1, act, ln):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)
    bias = torch.randn(N, generator=g) * 0.1
    ref = F.linear(A.double(), W.double(), bias.double())
    kw = {}
    if ln:
        R = torch.randn(M, N, generator=g)
        gamma, beta = 1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
        ref = F.layer_norm(ref + R.double(), (N,), gamma.double(), beta.double(), 1e-5)
        kw = dict(gamma=gamma, beta=beta, R=R)
    else:
        ref = {0: lambda x: x, 1: F.gelu, 2: F.relu, 3: F.silu}[act](ref)
    y_tc = eng.debug_gemm(A, W, bias, K1=K1, act=act, use_tc=True, **kw)
    y_cc = eng.debug_gemm(A, W, bias, K1=K1, act=act, use_tc=False, **kw)
    assert torch.isfinite(y_tc).all()
    assert _rel(y_cc, ref) < 5e-6, "CUDA-core kernel vs float64 reference"
    assert _rel(y_tc, ref) < 5e-6, "tcgen05 kernel vs float64 reference"
    if not ln and N % 8 == 0:
        # the production epilogue: split16 planes (TMA bulk stores on the CTA-pair kernels)
        y_sp = eng.debug_gemm(A, W, bias, K1=K1, act=act, use_tc=True, split_out=True)
        assert _rel(y_sp, ref) < 5e-6, "tcgen05 kernel, split16 output"


@pytest.mark.parametrize("M,ff", [(100, 1024), (128 * 3 + 5, 1024), (148 * 128 * 2 + 777, 1024), (1000, 512)])
def test_fused_ffn_block(eng, M, ff):
    """k_ffn_tc (hidden activations kept on the SM) vs float64 and vs the unfused operators."""
    d = 256
    g = torch.Generator().manual_seed(M + ff)
    X = torch.randn(M, d, generator=g)
    W1, b1 = torch.randn(ff, d, generator=g) / d ** 0.5, 0.1 * torch.randn(ff, generator=g)
    W2, b2 = torch.randn(d, ff, generator=g) / ff ** 0.5, 0.1 * torch.randn(d, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    Xd = X.double()
    hid = F.gelu(F.linear(Xd, W1.double(), b1.double()))
    ref = F.layer_norm(Xd + F.linear(hid, W2.double(), b2.double()), (d,), gamma.double(), beta.double(), 1e-5)
    ys = [eng.debug_ffn(X, W1, b1, W2, b2, gamma, beta, mode=m).cpu().double() for m in (0, 1, 2)]
    for name, y in zip(("cuda-core", "tc unfused", "tc fused"), ys):
        assert torch.isfinite(y).all(), name
        assert _rel(y, ref) < 5e-6, name
    assert _rel(ys[2], ys[1]) < 2e-6


def test_whole_path_tc_equals_cuda_core_path(built_lib):
    """The same 6-step sample through both GEMM paths agrees to fp32 re-association noise."""
    from mld_b200 import synth
    from mld_b200.engine import Engine, make_config
    eng = Engine(make_config(), 0)
    eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
    eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
    eng.finalize()
    eng.set_mean_std(*synth.mean_std())
    eng.set_timesteps(6)
    ctx, noise = synth.text_context(3, 77, seed=5), synth.init_noise(3, seed=6)
    lengths = [196, 64, 120]
    a = eng.sample(ctx, noise, lengths, want=("latents", "joints"))
    l_tc = eng.launch_count
    eng.set_option("gemm", "simt")
    b = eng.sample(ctx, noise, lengths, want=("latents", "joints"))
    assert _rel(a["latents"], b["latents"]) < 1e-4
    assert _rel(a["joints"], b["joints"]) < 1e-4
    assert l_tc > 0


def test_scheduling_options_do_not_change_results(built_lib):
    """Engine scheduling options only reorder independent work: the fused FFN pair launch (one
    persistent kernel running FFN1 -> FFN2 chains per CTA), the L2-sized producer/consumer chunking,
    the sequence chunking and the number of concurrent sub-batch branches must give bit-identical
    motions."""
    from mld_b200 import synth
    from mld_b200.engine import Engine, make_config
    eng = Engine(make_config(), 0)
    eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
    eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
    eng.finalize()
    eng.set_mean_std(*synth.mean_std())
    eng.set_timesteps(4)
    B = 300                                   # 600 sequences x 79 tokens = 371 m-tiles: > 2 waves + ragged tail
    ctx, noise = synth.text_context(B, 77, seed=15), synth.init_noise(B, seed=16)
    lengths = [196] * B
    fused = eng.sample(ctx, noise, lengths, want=("latents",))["latents"].clone()
    eng.set_option("ffn_fused", "0")          # the two-launch FFN: same math, hidden round-trips HBM
    base = eng.sample(ctx, noise, lengths, want=("latents",))["latents"].clone()
    assert _rel(fused, base) < 1e-5
    for name, value in (("branches", "1"), ("branches", "3"), ("ffn_pair", "1"), ("pair_chunk", "1"), ("chunk", "96")):
        eng.set_option(name, value)
        out = eng.sample(ctx, noise, lengths, want=("latents",))["latents"]
        assert torch.equal(out, base), f"option {name}={value} changed the result"
        eng.set_option(name, "0")
    # free-running per-lane chains (each lane holds the uncond + cond copies of its motions contiguously)
    eng.set_option("branches", "3")
    eng.set_option("lanes", "1")
    out = eng.sample(ctx, noise, lengths, want=("latents",))["latents"]
    assert torch.equal(out, base), "lanes changed the result"


@pytest.mark.skipif(os.environ.get("MLDB_EXPERIMENTAL") != "1",
                    reason="attn_tc.cu has not been validated on hardware yet (set MLDB_EXPERIMENTAL=1)")
def test_tc_attention_matches_mma(built_lib):
    """The experimental tcgen05 attention core (option attn_tc) against the product mma.sync kernel on
    the whole sampling path (ragged lengths exercise the key mask in the VAE decoder's fallback)."""
    from mld_b200 import synth
    from mld_b200.engine import Engine, make_config
    eng = Engine(make_config(), 0)
    eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
    eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
    eng.finalize()
    eng.set_mean_std(*synth.mean_std())
    eng.set_timesteps(4)
    ctx, noise = synth.text_context(5, 77, seed=25), synth.init_noise(5, seed=26)
    lengths = [196, 64, 120, 33, 196]
    a = eng.sample(ctx, noise, lengths, want=("latents", "joints"))
    a = {k: v.clone() for k, v in a.items()}
    eng.set_option("attn_tc", "1")
    b = eng.sample(ctx, noise, lengths, want=("latents", "joints"))
    assert torch.isfinite(b["latents"]).all()
    assert _rel(b["latents"], a["latents"]) < 1e-4
    assert _rel(b["joints"], a["joints"]) < 1e-4


def _attention_ref(qkv, nseq, L, heads, lengths=None):
    d = qkv.shape[1] // 3
    hd = d // heads
    q, k, v = (t.reshape(nseq, L, heads, hd).permute(0, 2, 1, 3).double() for t in qkv.split(d, dim=1))
    s = q @ k.transpose(-1, -2) / hd ** 0.5
    if lengths is not None:
        mask = torch.arange(L)[None, :] >= torch.as_tensor(lengths)[:, None]
        s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(nseq * L, d)


@pytest.mark.skipif(os.environ.get("MLDB_EXPERIMENTAL") != "1",
                    reason="debug hook + attn_tc.cu not validated on hardware yet (set MLDB_EXPERIMENTAL=1)")
@pytest.mark.parametrize("nseq,L,masked", [(3, 79, False), (5, 79, True), (2, 128, False), (4, 33, True), (300, 79, False)])
def test_attention_kernels_unit(eng, nseq, L, masked):
    """Attention cores in isolation against float64: CUDA-core, mma.sync (product) and tcgen05 (experimental)."""
    heads, hd = 4, 64
    g = torch.Generator().manual_seed(nseq * 131 + L)
    qkv = torch.randn(nseq * L, 3 * heads * hd, generator=g)
    lengths = [max(1, (7 * i + 5) % L) for i in range(nseq)] if masked else None
    ref = _attention_ref(qkv, nseq, L, heads, lengths)
    for mode, name in ((0, "cuda-core"), (1, "mma.sync"), (2, "tcgen05")):
        y = eng.debug_attention(qkv, nseq, L, heads, lengths, mode=mode).cpu().double()
        assert torch.isfinite(y).all(), name
        assert _rel(y, ref) < 5e-6, name
