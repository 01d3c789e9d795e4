"""CPU: carrier partition + the dibit all-gather on a world of 2 (gloo), and the host-side sync logic on the gathered streams."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tetraear_b200 import shard


@pytest.mark.parametrize("n,world", [(4096, 8), (4096, 1), (7, 2), (5, 8), (0, 2), (96, 8), (13, 4)])
def test_partition_covers_all_carriers_once(n, world):
    seen = []
    for r in range(world):
        first, count = shard.partition(n, world, r)
        assert count >= 0 and first == len(seen)
        seen.extend(range(first, first + count))
        for c in range(first, first + count):
            assert shard.owner_of(c, n, world) == r
    assert seen == list(range(n))
    sizes = [shard.partition(n, world, r)[1] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_streams(n_carriers, cap):
    rng = np.random.default_rng(99)
    dib = rng.integers(0, 4, size=(n_carriers, cap), dtype=np.uint8)
    nd = rng.integers(cap // 2, cap + 1, size=n_carriers).astype(np.int32)
    for c in range(n_carriers):
        dib[c, nd[c]:] = 0
    return dib, nd


def _worker(rank, world, port, n_carriers, cap, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dib, nd = _fake_streams(n_carriers, cap)
        first, count = shard.partition(n_carriers, world, rank)
        d_all, n_all = shard.gather_dibits(torch.from_numpy(dib[first:first + count].copy()),
                                           torch.from_numpy(nd[first:first + count].copy()), n_carriers)
        good = np.array_equal(d_all.numpy(), dib) and np.array_equal(n_all.numpy(), nd)
        t = shard.max_over_ranks(float(rank + 1))
        good = good and t == float(world)
        if n_carriers % world == 0:                      # the single-collective variant
            ps = shard.PackedStreams(n_carriers, cap)
            ps.dibits[:, :cap] = torch.from_numpy(dib[first:first + count].copy())
            ps.n_dibits[:] = torch.from_numpy(nd[first:first + count].copy())
            d2, n2 = ps.gather()
            good = good and np.array_equal(d2.reshape(n_carriers, -1).numpy()[:, :cap], dib)
            good = good and np.array_equal(n2.reshape(-1).numpy(), nd)
        ok[rank] = 1 if good else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_carriers", [8, 7])
def test_gather_dibits_world2_gloo(n_carriers):
    world, cap = 2, 301
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = _free_port()
    procs = [mp.get_context("spawn").Process(target=_worker, args=(r, world, port, n_carriers, cap, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


def test_transport_packing_round_trip():
    rng = np.random.default_rng(3)
    d = torch.from_numpy(rng.integers(0, 4, size=64, dtype=np.uint8))
    p = shard._pack_cpu(d)
    assert p.shape == (16,) and int(p[0]) == int(d[0]) | int(d[1]) << 2 | int(d[2]) << 4 | int(d[3]) << 6
    assert torch.equal(shard._unpack_cpu(p), d)


def test_gather_is_identity_without_process_group():
    d = torch.zeros((3, 5), dtype=torch.uint8)
    n = torch.zeros(3, dtype=torch.int32)
    a, b = shard.gather_dibits(d, n, 3)
    assert a is d and b is n
    assert shard.max_over_ranks(2.5) == 2.5
