"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: weight broadcast from rank 0, the
reference's rank-strided work split, per-rank seeds, max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from gpt_image_edit_b200 import distributed as D

    w, r, _ = D.init_from_env(backend="gloo")
    assert (w, r) == (world, rank)
    g = torch.Generator().manual_seed(0)
    weights = [torch.randn(64, 32, generator=g).bfloat16(), torch.randn(7, generator=g).bfloat16()]
    if rank != 0:
        for t in weights:
            t.zero_()
    D.broadcast_weights(weights, src=0)
    items = list(range(11))
    mine = D.shard(items, rank, world)
    t_max = D.max_over_ranks(10.0 + rank)
    D.barrier()
    torch.save(dict(weights=weights, items=mine, seed=D.rank_seed(42, rank), t_max=t_max), f"{out_dir}/r{rank}.pt")
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_broadcast_shard_and_timing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    for x, y in zip(a["weights"], b["weights"]):
        assert torch.equal(x, y) and x.abs().sum() > 0          # rank 1 received rank 0's weights
    assert a["items"] == [0, 2, 4, 6, 8, 10] and b["items"] == [1, 3, 5, 7, 9]   # reference striding, disjoint cover
    assert (a["seed"], b["seed"]) == (42, 43)
    assert a["t_max"] == b["t_max"] == 11.0


def test_single_process_is_a_noop():
    from gpt_image_edit_b200 import distributed as D

    os.environ.pop("WORLD_SIZE", None)
    assert D.env_world()[0] == 1
    D.broadcast_weights([torch.ones(2)])
    assert D.max_over_ranks(3.5) == 3.5 and D.shard([1, 2, 3], 0, 1) == [1, 2, 3]
