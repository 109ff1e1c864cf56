"""Host-side logic of the multi-GPU Monte-Carlo driver on CPU: seeds are dealt round-robin, partial
sums are all-reduced once (gloo, world_size 2), and the result equals the serial sum.  The
realisation itself is a deterministic stand-in here; the CUDA pipelines are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectra_without_windows_b200 import mc_driver


def fake_realisation(seed, pre):
    rng = np.random.RandomState(seed)
    F = torch.from_numpy(rng.normal(size=(6, 6)))
    q = torch.from_numpy(rng.normal(size=6)) + (0.0 if pre is None else float(pre))
    return F, q


def zeros():
    return torch.zeros((6, 6), dtype=torch.float64), torch.zeros(6, dtype=torch.float64)


def test_shard_is_a_partition():
    seeds = list(range(1, 101))
    for world in (1, 2, 3, 8):
        parts = [mc_driver.shard(seeds, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == seeds
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_serial_run_with_prefetch():
    seeds = range(1, 8)
    (F, q), n = mc_driver.run_mc(seeds, fake_realisation, zeros, prefetch_fn=lambda s: 0.5 * s, workers=3)
    Fr, qr = zeros()
    for s in seeds:
        f, g = fake_realisation(s, 0.5 * s)
        Fr += f
        qr += g
    assert n == 7 and torch.equal(F, Fr) and torch.equal(q, qr)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    (F, q), n = mc_driver.run_mc(range(1, 12), fake_realisation, zeros, rank, world, dist)
    if rank == 0:
        torch.save({"F": F, "q": q, "n": n}, out)
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_serial(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    (F, q), n = mc_driver.run_mc(range(1, 12), fake_realisation, zeros)
    assert got["n"] == 6 and n == 11
    assert torch.allclose(got["F"], F, rtol=0, atol=1e-13) and torch.allclose(got["q"], q, rtol=0, atol=1e-13)
