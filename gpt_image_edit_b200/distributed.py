"""Multi-GPU plumbing: one process per GPU, batch-sharded replicas.

The path shards by independent edit requests exactly as the reference's eval drivers do
(univa/eval/gedit/step1_gen_samples.py:82-92 init, :239 `inference_list[rank::world_size]`, :33-42
`seed + rank`): every rank holds a full replica, takes items rank, rank+N, ..., and never talks to the
others inside the sampling loop.  The only collective is ONE broadcast of the weights from rank 0 at
load (NCCL over NVLink on the GPU box; gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str | None = None, device: torch.device | None = None):
    """Initialise torch.distributed from torchrun's env (no-op for a single process). Returns (world, rank, local)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world, rank, local


def shard(items, rank: int, world: int):
    """The reference's striding: rank r takes items r, r+N, r+2N, ..."""
    return list(items)[rank::world]


def rank_seed(seed: int, rank: int) -> int:
    return seed + rank


def broadcast_weights(tensors, src: int = 0):
    """Broadcast every tensor in place from `src` (weights drawn/loaded once on rank 0)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU timings are the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
