"""2-rank diagnostic: where does the ZeRO-2 update diverge from the single-rank one?"""
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    from gpt_image_edit_b200 import distributed as D
    from univa.training.configuration_denoise import from_mapping
    from univa.training.synthetic_data import SyntheticEditDataset, collate
    import train_denoiser as td
    from gpt_image_edit_b200.training import Stage2Trainer, ShardedAdamW

    world, rank, local = D.env_world()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    D.init_from_env(device=dev)
    conf = from_mapping(dict(
        training_config=dict(seed=5, learning_rate=1e-4, adam_beta2=0.99, adam_weight_decay=0.0, discrete_timestep=False,
                             mask_weight_type="log", max_grad_norm=1e9),
        model_config=dict(synthetic=True, small=True, with_tune_mlp2=True, joint_ref_feature=True,
                          flux_train_layer_idx=list(range(57))),
        dataset_config=dict(dataset_type="synthetic", batch_size=1, height=256, width=256)))
    model, vae, pipe, empty = td.build_models(conf, dev)
    D.broadcast_weights(list(model.denoise_tower.denoiser._store.values()) +
                        list(model.denoise_tower.denoise_projector.state_dict().values()))
    tr = Stage2Trainer(model, vae, pipe, conf.training_config, conf.model_config, empty)
    data = SyntheticEditDataset(256, 256, seed=1)
    w0 = [p.storage.clone() for p in tr.params]
    # run the step but stop before opt.step: replicate Stage2Trainer.step's tail by hand
    orig_step = tr.opt.step
    captured = {}

    def fake_step(lr=None):
        torch.cuda.synchronize()
        captured["local"] = [p.grad.clone() for p in tr.params]
        captured["shard"] = [bk.shard_grad.clone() if bk is not None else None for bk in tr.opt.buckets]
        return orig_step(lr)
    tr.opt.step = fake_step
    tr.gen = torch.Generator(device=dev).manual_seed(100)
    out = tr.step(collate([data[0]]))
    torch.cuda.synchronize()
    # (a) local grads equal across ranks?
    for p, g in zip(tr.params, captured["local"]):
        ref = g.clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(ref, g):
            print(f"[rank {rank}] LOCAL GRAD differs from rank 0: {p.name} max|d|={float((ref - g).abs().max()):.3e} of {float(g.abs().max()):.3e}")
    # (b) shard == 2 * local slice?
    for b, bk in enumerate(tr.opt.buckets):
        if bk is None:
            continue
        n = bk.size // world
        # rebuild the local flat from the captured per-param grads
        flat = torch.zeros(bk.size, device=dev)
        for p in bk.params:
            i = tr.params.index(p)
            flat[p.offset:p.offset + p.storage.numel()] = captured["local"][i].reshape(-1)
        want = 2 * flat[rank * n:(rank + 1) * n]
        got = captured["shard"][b]
        if not torch.equal(want, got):
            d = (want - got).abs()
            print(f"[rank {rank}] bucket {b}: shard != 2*local: max|d|={float(d.max()):.3e} nnz={int((d > 0).sum())} of {n}")
    # (c) emulate the single-rank update on this rank and compare with the stored weights
    for p, g, w in zip(tr.params, captured["local"], w0):
        g32 = g.reshape(-1)
        p32 = w.float().reshape(-1)
        m = 0.1 * g32
        v = 0.01 * g32 * g32
        upd = p32 - (1e-4 / 0.1) * m / (v.sqrt() / (0.01 ** 0.5) + 1e-8)
        exp = upd.to(torch.bfloat16).view(p.storage.shape)
        if not torch.equal(exp, p.storage):
            d = (exp.float() - p.storage.float()).abs()
            print(f"[rank {rank}] {p.name}: weights != emulated single-rank update: nnz={int((d > 0).sum())} of {d.numel()} max={float(d.max()):.3e}")
    dist.barrier()
    if rank == 0:
        print("diag done", out["loss"].item(), out["grad_norm"].item())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
