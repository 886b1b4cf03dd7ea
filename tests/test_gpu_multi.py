"""Multi-GPU parity (pytest -m gpu, skipped on a box with fewer than 2 GPUs): two processes, one per GPU, the
real engine and the real collective (ncclAllGather through the C ABI) reproduce the single-GPU result
bit for bit ('the gathered result equals the single-GPU result')."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

B, S, STEPS = 6, 77, 4
LENGTHS = [196, 64, 120, 33, 196, 100]


def _engine(device):
    from mld_b200 import synth
    from mld_b200.engine import Engine, make_config
    eng = Engine(make_config(), device)
    eng.load_state_dict(synth.denoiser_state_dict(1234), "denoiser.")
    eng.load_state_dict(synth.mld_vae_state_dict(4321), "vae.")
    eng.finalize()
    eng.set_mean_std(*synth.mean_std())
    eng.set_timesteps(STEPS)
    return eng


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from mld_b200 import synth
    from mld_b200.distributed import sample_sharded, sample_sharded_engine
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    eng = _engine(rank)
    assert eng.comm_init() == (world, rank)
    ctx, noise = synth.text_context(B, S, seed=401), synth.init_noise(B, seed=402)
    # the C-ABI collective (product path), twice: alternate buffers, the gather overlaps the next batch
    j1 = sample_sharded_engine(eng, ctx, noise, LENGTHS)
    j2 = sample_sharded_engine(eng, ctx, noise, LENGTHS)
    # the generic torch.distributed path (ragged splits) on the same engine
    j3 = sample_sharded(lambda c, z, ln: eng.sample(c, z, ln, want=("joints",))["joints"], ctx, noise, LENGTHS)
    torch.cuda.synchronize()
    assert torch.equal(j1, j2)
    torch.save({"abi": j1.cpu(), "torch": j3.cpu()}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sample_equals_single_gpu(tmp_path, built_lib):
    from mld_b200 import synth
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "gathered.pt")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    eng = _engine(0)
    c, z = synth.text_context(B, S, seed=401), synth.init_noise(B, seed=402)
    want = eng.sample(c, z, LENGTHS, want=("joints",))["joints"].cpu()
    for r in range(2):
        got = torch.load(f"{out_path}.{r}")
        assert torch.equal(got["abi"], want), f"rank {r}: C-ABI gather differs from the single-GPU result"
        assert torch.equal(got["torch"], want), f"rank {r}: torch.distributed gather differs"


def test_single_rank_gather_is_identity(built_lib):
    """Without a communicator mldb_sample_gather degenerates to mldb_sample (world = 1)."""
    from mld_b200 import synth
    eng = _engine(0)
    c, z = synth.text_context(B, S, seed=401), synth.init_noise(B, seed=402)
    a = eng.sample(c, z, LENGTHS, want=("joints",))["joints"]
    b = eng.sample_gather(c, z, LENGTHS)
    assert torch.equal(a, b)
