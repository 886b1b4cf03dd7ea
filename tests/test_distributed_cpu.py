"""CPU, world_size 2, gloo: the batch-sharding host logic and the single all-gather reproduce
the single-rank result.  The per-rank sampler is a stand-in function -
the collective plumbing is what is under test here; the GPU tests cover the kernels."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mld_b200.distributed import sample_sharded, shard_cfg_condition, shard_range


def _fake_sampler(cond, noise, lengths):
    # deterministic per-motion function of (uncond row, cond row, noise row); pads to local max
    b = noise.shape[0]
    T = max(lengths)
    out = torch.zeros(b, T, 3)
    for i in range(b):
        v = cond[i].sum() * 0.5 + cond[b + i].sum() + noise[i].sum()
        out[i, : lengths[i]] = v + torch.arange(lengths[i])[:, None] * torch.tensor([1.0, 2.0, 3.0])
    return out


def _worker(rank, world, port, B, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    cond = torch.randn(2 * B, 4, 8, generator=g)
    noise = torch.randn(B, 1, 16, generator=g)
    lengths = (torch.randint(3, 12, (B,), generator=g)).tolist()
    out = sample_sharded(_fake_sampler, cond, noise, lengths, cfg_on=True)
    if rank == 0:
        torch.save(out, out_path)          # a file, not a queue: nothing to drain while a rank exits
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(B, tmp_path):
    ctx = mp.get_context("spawn")
    out_path = str(tmp_path / "gathered.pt")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    out = torch.load(out_path)
    g = torch.Generator().manual_seed(0)
    cond = torch.randn(2 * B, 4, 8, generator=g)
    noise = torch.randn(B, 1, 16, generator=g)
    lengths = (torch.randint(3, 12, (B,), generator=g)).tolist()
    want = _fake_sampler(cond, noise, lengths)
    assert out.shape == want.shape and torch.equal(out, want)      # bit-exact


def test_shard_ranges_cover_batch():
    for B in (1, 7, 256, 1024):
        for w in (1, 2, 4, 8):
            r = [shard_range(B, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == B
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
    c = torch.arange(8)[:, None]
    assert shard_cfg_condition(c, 4, 1, 3, True).flatten().tolist() == [1, 2, 5, 6]


def test_two_rank_gather_even(tmp_path):
    _run(8, tmp_path)


def test_two_rank_gather_uneven(tmp_path):
    _run(7, tmp_path)
