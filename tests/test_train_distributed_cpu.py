"""world_size-2 `gloo` test (CPU) of the ZeRO-2 optimiser's host logic (training.ShardedAdamW): bucket layout and padding,
reduce-scatter of per-bucket gradients, per-rank fp32 partitions, global-norm clipping across partitions, all-gather of
the updated bf16 weights into the (fused, sliced) model storage.  The CUDA kernels are replaced by the oracle's torch
arithmetic (oracle.train_oracle.TorchMath) — this test is about the partitioning, the kernels have their own GPU tests —
and the result on every rank must equal the un-partitioned reference update (torch.optim.AdamW + clip_grad_norm_ on the
rank-averaged gradients).  Reference: scripts/accelerate_configs/zero2.json, train_denoiser.py:596-602, 1172-1181."""
import os
import socket

import pytest

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


SHAPES = [(96, 64), (96,), (64, 64), (128,), (33, 7), (5,)]       # odd sizes: padding inside and at the end of buckets
BUCKETS = [0, 0, 1, 1, 3, 3]                                        # bucket 2 is empty (a frozen block)
HP = dict(lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.02, max_grad_norm=0.5)


def _make(seed=0):
    from gpt_image_edit_b200.training import Param
    g = torch.Generator().manual_seed(seed)
    big = torch.randn(400, 64, generator=g).bfloat16()              # params are views of larger storages, like the fused tensors
    stor = [big[:96], torch.randn(96, generator=g).bfloat16(), big[100:164], torch.randn(128, generator=g).bfloat16(),
            torch.randn(33, 7, generator=g).bfloat16(), torch.randn(5, generator=g).bfloat16()]
    return [Param(f"p{i}", s, None, b) for i, (s, b) in enumerate(zip(stor, BUCKETS))]


def _grads(rank, step):
    g = torch.Generator().manual_seed(1000 * (step + 1) + rank)
    return [torch.randn(s, generator=g) * (1 + rank) for s in SHAPES]


def _worker(rank, world, port, out_dir):
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo")
    from gpt_image_edit_b200.training import ShardedAdamW
    from oracle.train_oracle import TorchMath

    params = _make()
    opt = ShardedAdamW(params, math=TorchMath, lr=HP["lr"], betas=HP["betas"], eps=HP["eps"], weight_decay=HP["weight_decay"],
                       max_grad_norm=HP["max_grad_norm"])
    assert opt.world == world and opt.buckets[2] is None
    for bk in (b for b in opt.buckets if b is not None):
        assert bk.size % (world * 64) == 0 and bk.p32.numel() == bk.size // world
    norms = []
    for step in range(3):
        for p, g in zip(params, _grads(rank, step)):
            p.grad.copy_(g)
        opt.reduce_all()
        norms.append(float(opt.step()))
    torch.save(dict(weights=[p.storage.clone() for p in params], norms=norms,
                    state_numel=sum(b.p32.numel() for b in opt.buckets if b is not None)), f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


def test_zero2_two_ranks_equal_the_unpartitioned_update(tmp_path):
    from oracle.train_oracle import reference_step

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    a, b = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    for x, y in zip(a["weights"], b["weights"]):
        assert torch.equal(x, y)                                     # every rank holds the same updated weights
    weights = [p.storage.clone() for p in _make()]
    state, ref_norms = None, []
    for step in range(3):
        weights, norm, state = reference_step(weights, [_grads(r, step) for r in range(world)], state, **HP)
        ref_norms.append(float(norm))
    for w_ref, w in zip(weights, a["weights"]):
        diff = (w_ref.float() - w.float()).abs().max().item()
        assert diff <= 2 ** -7 * max(1.0, w_ref.float().abs().max().item()), diff   # at most one bf16 ulp (fp32 summation order)
        assert (w_ref != w).float().mean().item() < 0.02
    for n, r in zip(a["norms"], ref_norms):
        assert abs(n - r) < 1e-4 * r
    total = sum(torch.Size(s).numel() for s in SHAPES)
    assert a["state_numel"] < 0.75 * total + 3 * 64                 # each rank holds about half of the optimizer state


def test_single_rank_sharded_adamw_is_plain_adamw():
    from gpt_image_edit_b200.training import ShardedAdamW
    from oracle.train_oracle import TorchMath, reference_step

    params = _make(seed=3)
    before = [p.storage.clone() for p in params]
    opt = ShardedAdamW(params, math=TorchMath, lr=HP["lr"], betas=HP["betas"], eps=HP["eps"], weight_decay=HP["weight_decay"],
                       max_grad_norm=HP["max_grad_norm"])
    state = None
    weights = before
    for step in range(2):
        gs = _grads(0, step)
        for p, g in zip(params, gs):
            p.grad.copy_(g)
        opt.reduce_all()
        opt.step()
        weights, _, state = reference_step(weights, [gs], state, **HP)
    for w_ref, p in zip(weights, params):
        assert (w_ref.float() - p.storage.float()).abs().max().item() <= 2 ** -7 * max(1.0, w_ref.float().abs().max().item())
    opt.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0 for p in params)


def test_trainable_set_names_follow_the_references_contract():
    """get_trainable_params / check_param_is_in_components (train_denoiser.py:71-119) as restated in training.py against
    the outputs of the reference's own function sources (tests/golden/host_ref.pt, made by make_host_ref_golden.py)."""
    from pathlib import Path

    from gpt_image_edit_b200 import training as tr

    h = torch.load(Path(__file__).parent / "golden" / "host_ref.pt", weights_only=False)["host"]
    # "default", "both_branches", "some_layers": outputs of the reference's own function source for three argument sets
    assert tr.get_trainable_params() == h["components"]["default"]
    assert tr.get_trainable_params(only_img_branch=False) == h["components"]["both_branches"]
    some = h["components"]["some_layers"]
    layers = sorted({int(c.split("transformer_blocks.")[1].split(".")[0]) + (19 if "single_" in c else 0) for c in some})
    assert tr.get_trainable_params(layers) == some
    for mode, want in h["probe_result"].items():
        assert [tr.check_param_is_in_components(n, h["components"][mode]) for n in h["probe"]] == want
    comps = tr.get_trainable_params([0, 20], 19, True)
    assert comps[0] == "denoise_tower.denoiser.transformer_blocks.0.attn.norm_q"
    assert "denoise_tower.denoiser.single_transformer_blocks.1.norm.linear" in comps and len(comps) == 7 + 6


def _resume_worker(rank, world, port, out_dir):
    """3 steps straight through  vs  2 steps, state_dict -> a FRESH optimizer over the INITIAL weights, load, 1 more step."""
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("gloo")
    from gpt_image_edit_b200.training import ShardedAdamW
    from oracle.train_oracle import TorchMath

    def run(opt, params, steps):
        for step in steps:
            for p, g in zip(params, _grads(rank, step)):
                p.grad.copy_(g)
            opt.reduce_all()
            opt.step()

    kw = dict(math=TorchMath, lr=HP["lr"], betas=HP["betas"], eps=HP["eps"], weight_decay=HP["weight_decay"],
              max_grad_norm=HP["max_grad_norm"])
    straight = _make()
    run(ShardedAdamW(straight, **kw), straight, range(3))
    first = _make()
    opt = ShardedAdamW(first, **kw)
    run(opt, first, range(2))
    torch.save(opt.state_dict(), f"{out_dir}/opt{rank}.pt")
    after_two = [p.storage.clone() for p in first]
    resumed = _make()                                                # the weights a restarted process starts from
    opt2 = ShardedAdamW(resumed, **kw)
    opt2.load_state_dict(torch.load(f"{out_dir}/opt{rank}.pt"))
    restored = all(torch.equal(a, p.storage) for a, p in zip(after_two, resumed))
    run(opt2, resumed, [2])
    same = all(torch.equal(a.storage, b.storage) for a, b in zip(straight, resumed))
    torch.save(dict(restored=restored, same=same, steps=opt2.step_count), f"{out_dir}/res{rank}.pt")
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_optimizer_state_restores_the_models_weights_and_continues_bit_exactly(tmp_path, world):
    """Resume (train_denoiser.py:769 accelerator.load_state): loading a rank's optimizer partition must also bring back the
    bf16 weights the model computes with — a restarted process holds the INITIAL weights — and the next step must equal the
    uninterrupted run's."""
    if world == 1:
        _resume_worker(0, 1, 0, str(tmp_path))
    else:
        mp.spawn(_resume_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = torch.load(tmp_path / f"res{r}.pt")
        assert res["restored"] and res["same"] and res["steps"] == 3, res


def test_optimizer_state_of_another_partitioning_is_refused():
    from gpt_image_edit_b200 import _lib
    from gpt_image_edit_b200.training import ShardedAdamW
    from oracle.train_oracle import TorchMath

    opt = ShardedAdamW(_make(), math=TorchMath)
    sd = opt.state_dict()
    with pytest.raises(_lib.B2FError):
        opt.load_state_dict({**sd, "world": 2})
    with pytest.raises(_lib.B2FError):
        opt.load_state_dict({**sd, "buckets": sd["buckets"][:-1]})
