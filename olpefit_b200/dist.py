"""Multi-GPU plumbing: walkers shard across ranks, nothing else does.

The reference's ranks never exchange data (the only MPI call is a barrier, apf_step2.py:338),
so there is no data-path collective.  torch.distributed (NCCL on GPUs, gloo in CPU tests) is
used for exactly three things: summing the try/accept counters, taking the global minimum of
tries for the stop rule (apf_step2.py:300, globalised so all chain files keep equal length as
apf_step3.py:183 requires), summing Gelman-Rubin moments; plus an optional final gather.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_ids(n_walkers_total: int, rank: int, world: int):
    """Interleaved sharding: rank r owns global walker ids r, r+world, ...  Returns
    (id_base, id_stride, n_local).  A walker's random stream is keyed by its global id, so the
    chains are identical for any world size."""
    n_local = (n_walkers_total - rank + world - 1) // world if n_walkers_total > rank else 0
    return rank, world, n_local


def allreduce_stats(tries: torch.Tensor, accepts: torch.Tensor, min_tries: torch.Tensor):
    """Global counters: sum of tries/accepts over all walkers of all ranks, min of tries."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        both = torch.cat([tries.reshape(-1), accepts.reshape(-1)])
        dist.all_reduce(both, op=dist.ReduceOp.SUM)
        n = tries.numel()
        tries, accepts = both[:n].reshape(tries.shape), both[n:].reshape(accepts.shape)
        min_tries = min_tries.clone()
        dist.all_reduce(min_tries, op=dist.ReduceOp.MIN)
    return tries, accepts, min_tries


def allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max(t: torch.Tensor) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_chains(local: torch.Tensor, rank: int, world: int):
    """Final chain gather to rank 0: local [rows, W_local, C] -> list of per-rank tensors on rank 0
    (interleave with ``merge_interleaved``).  Shards may differ by one walker, so sizes are
    exchanged first and the payload is padded."""
    if world == 1 or not dist.is_initialized():
        return [local]
    n = torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    wmax = int(max(int(s) for s in sizes))
    pad = torch.zeros((local.shape[0], wmax, local.shape[2]), dtype=local.dtype, device=local.device)
    pad[:, :local.shape[1]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, parts, dst=0)
    if rank != 0:
        return None
    return [p[:, :int(s)] for p, s in zip(parts, sizes)]


def merge_interleaved(parts):
    """Inverse of ``shard_ids``: parts[r] holds global walkers r, r+world, ..."""
    world = len(parts)
    total = sum(p.shape[1] for p in parts)
    out = torch.empty((parts[0].shape[0], total, parts[0].shape[2]), dtype=parts[0].dtype,
                      device=parts[0].device)
    for r, p in enumerate(parts):
        out[:, r::world] = p
    return out
