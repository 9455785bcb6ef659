"""CPU: host-side multi-GPU logic — slice bounds and the closing all_gather — with world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastamericanoptionpricing_b200 import workloads as wl
from fastamericanoptionpricing_b200.sharding import gather_interleaved, gather_slices, interleaved_rows, shard_bounds


def test_shard_bounds_cover_and_balance():
    for N in (0, 1, 7, 1000, 1_000_003):
        for W in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(N, W, r) for r in range(W)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_rows_cover():
    for G in (0, 1, 7, 1000, 100_003):
        for W in (1, 2, 3, 8):
            rows = sorted(i for r in range(W) for i in range(G)[interleaved_rows(G, W, r)])
            assert rows == list(range(G))
            sizes = [len(range(G)[interleaved_rows(G, W, r)]) for r in range(W)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(N, world, rank)
    full = torch.arange(N, dtype=torch.float64) * 1.5 + 0.25          # stands in for the price vector
    local = full[lo:hi].clone()
    got = gather_slices(local, N, world)
    ok = torch.equal(got, full)
    # row-wise cut (ladder_batch_sharded: scenarios x strikes): N rows of 3
    mat = torch.arange(3 * N, dtype=torch.float64).reshape(N, 3)
    ok = ok and torch.equal(gather_slices(mat[lo:hi].reshape(-1).clone(), 3 * N, world, unit=3).reshape(N, 3), mat)
    # round-robin deal (ladder_batch_sharded / bench.py's C5: scenario g -> rank g mod W)
    sl = interleaved_rows(N, world, rank)
    ok = ok and torch.equal(gather_interleaved(mat[sl].clone(), N, world, unit=3).reshape(N, 3), mat)
    # an exact, order-independent checksum (integer micro-units), equal for every world size, as bench.py's C5 line reports
    chk = torch.round(mat[sl] * 1e6).to(torch.int64).sum().reshape(1)
    dist.all_reduce(chk)
    ok = ok and chk.item() == int(torch.round(mat * 1e6).to(torch.int64).sum())
    # every rank reduces its step time with MAX, as bench.py does
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = bool(ok) and t.item() == 10.0 + world - 1
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [1001, 4096])
def test_gather_slices_world2_gloo(N):
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), N, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


def test_workload_generators_are_pinned():
    b = wl.c2_batch(64)
    # same draw order as SURVEY App. C's generator (first option of the N=64 batch prices to 61.1267517361... upstream)
    assert b["is_put"][:4].tolist() == (np.random.default_rng(wl.C2_SEED).random(64) < 0.5)[:4].tolist()
    assert abs(b["r"][0] - np.random.default_rng(wl.C2_SEED).uniform(0, 1, 128)[64] * 0.099 - 0.001) < 1e-15
    b3 = wl.c3_batch(12)
    assert b3["category"].tolist() == [0, 1, 2] * 4
    assert np.all(b3["sigma"][0::3] <= 0.10) and np.all(b3["T"][1::3] >= 5.0) and np.all(b3["q"][2::3] > b3["r"][2::3])
    s = wl.c5_surface(0, 5)
    assert s["K"][0] == 50.0 and s["sigma"].tolist() == list(np.linspace(0.10, 0.60, 100)[:5])
    last = wl.c5_surface(100_000_000 - 1, 1)
    assert last["K"][0] == 150.0 and last["T"][0] == 3.0 and last["sigma"][0] == 0.6
