"""Multi-rank host logic on CPU: two gloo processes exercise the sharding map, the statistics
all-reduces and the final gather of olpefit_b200/dist.py -- the only things that cross ranks on
this path (the reference's ranks share nothing but a barrier, apf_step2.py:338)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from olpefit_b200 import dist
    r, lr, w = dist.init(backend="gloo")
    assert (r, w) == (rank, world)
    base, stride, n_local = dist.shard_ids(total, rank, world)
    ids = base + stride * np.arange(n_local)
    # a fake per-walker state that depends on the GLOBAL id only, like the Philox-keyed chains
    P, rows = 16, 5
    tries = torch.tensor(np.stack([(g * 7 + np.arange(P)) % 11 + 3 for g in ids]), dtype=torch.int64)
    accepts = tries // 2
    t, a, mn = dist.allreduce_stats(tries.sum(0), accepts.sum(0), tries.min())
    chain = torch.tensor(np.stack([np.full((rows, P + 1), float(g)) + np.arange(rows)[:, None] for g in ids], axis=1))
    parts = dist.gather_chains(chain, rank, world)
    mom = dist.allreduce_sum(torch.tensor([float(ids.sum()), float((ids ** 2).sum())], dtype=torch.float64))
    mx = dist.allreduce_max(torch.tensor([float(rank)]))
    dist.barrier()
    if rank == 0:
        merged = dist.merge_interleaved(parts)
        np.savez(os.path.join(out_dir, "r0.npz"), tries=t.numpy(), accepts=a.numpy(), mn=int(mn),
                 merged=merged.numpy(), mom=mom.numpy(), mx=float(mx))
    else:
        assert parts is None
    td.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_two_rank_statistics_and_gather(tmp_path, total):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    z = np.load(str(tmp_path / "r0.npz"))
    P = 16
    all_tries = np.stack([(g * 7 + np.arange(P)) % 11 + 3 for g in range(total)])
    assert np.array_equal(z["tries"], all_tries.sum(0))
    assert np.array_equal(z["accepts"], (all_tries // 2).sum(0))
    assert z["mn"] == all_tries.min()
    # gathered chains come back in global walker order, whatever the shard sizes (7 walkers: 4 + 3)
    assert z["merged"].shape == (5, total, P + 1)
    for g in range(total):
        assert np.all(z["merged"][:, g, 0] == g + np.arange(5))
    ids = np.arange(total)
    assert np.array_equal(z["mom"], [ids.sum(), (ids ** 2).sum()]) and z["mx"] == 1.0


def test_shard_ids_partition_every_walker_once():
    from olpefit_b200 import dist
    for total in (1, 5, 64, 65536, 1000003):
        for world in (1, 2, 4, 8):
            seen = 0
            for r in range(world):
                base, stride, n = dist.shard_ids(total, r, world)
                assert base == r and stride == world
                if n:
                    assert base + stride * (n - 1) < total <= base + stride * n + (world - 1 - r) + r
                seen += n
            assert seen == total
    assert dist.env_world() == (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
                                int(os.environ.get("WORLD_SIZE", 1)))
