"""NCCL check of the ZeRO-2 training step (run on a GPU box, once per world size, then compare):

    python scripts/check_zero2_nccl.py                                   # world 1  -> gpurun_out/zero2_w1.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_zero2_nccl.py
    python scripts/check_zero2_nccl.py --compare 1 2

Every rank draws the SAME synthetic batch and noise, so the rank-averaged gradient equals the single-rank gradient
exactly (x + x = 2x and 2x / 2 = x are exact in fp32): after three steps the trained weights of the 2-rank ZeRO-2 run
must be BIT-IDENTICAL to those of the 1-rank run (sha256 of every trainable tensor), every rank must hold the same
weights, and the optimizer state must be partitioned 1/world per rank.
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
                             mask_weight_type="log", max_grad_norm=1e9),   # no clipping: the clip coefficient depends on the summation order of the partitions
        model_config=dict(synthetic=True, small=True, with_tune_mlp2=True, joint_ref_feature=True,
                                               flux_train_layer_idx=list(range(57))),
        dataset_config=dict(dataset_type="synthetic", batch_size=1, height=256, width=256)))
    model, vae, pipe, empty = td.build_models(conf, dev)
    D.broadcast_weights(list(model.denoise_tower.denoiser._store.values()) +
                        list(model.denoise_tower.denoise_projector.state_dict().values()))
    tr = Stage2Trainer(model, vae, pipe, conf.training_config, conf.model_config, empty,
                       overlap_comm=os.environ.get('ZERO2_NO_OVERLAP') is None)
    data = SyntheticEditDataset(256, 256, seed=1)
    if os.environ.get('ZERO2_FORCE_BLOCKS'):
        tr.graph.on_block_done = lambda b: None
    tag = os.environ.get('ZERO2_TAG', '')
    res = {"world": world}
    tr.trace = []
    import hashlib
    def all_sha():
        h = hashlib.sha256()
        for k, t in model.denoise_tower.denoiser._store.items():
            h.update(t.contiguous().view(torch.int16).cpu().numpy().tobytes())
        return h.hexdigest()
    res['store_sha'] = [all_sha()]
    for step in range(int(os.environ.get('ZERO2_STEPS', '3'))):
        tr.gen = torch.Generator(device=dev).manual_seed(100 + step)          # same noise / sigma on every rank
        out = tr.step(collate([data[step]]))
        res.setdefault("loss", []).append(out["loss"].item())
        res.setdefault("grad_norm", []).append(out["grad_norm"].item())
        res['store_sha'].append(all_sha())
    res['trace'] = tr.trace
    # all ranks hold identical weights
    if world > 1:
        for p in tr.params:
            ref = p.storage.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(ref, p.storage), p.name
    # optimizer state is partitioned
    state = sum(b.p32.numel() for b in tr.opt.buckets if b is not None)
    total = sum(b.size for b in tr.opt.buckets if b is not None)
    assert state * world == total
    if world > 1:
        dist.barrier()
    import hashlib
    res["sha256"] = {p.name: hashlib.sha256(p.storage.contiguous().view(torch.int16).cpu().numpy().tobytes()).hexdigest()
                     for p in tr.params}
    res["optimizer_state_fraction"] = state / total
    res["sums"] = {p.name: [p.storage.double().sum().item(), p.storage.double().abs().sum().item()] for p in tr.params}
    res["grad_sums"] = {p.name: [p.grad.double().sum().item(), p.grad.double().abs().sum().item()] for p in tr.params}
    torch.cuda.synchronize()
    t0 = time.time()
    for step in range(5):
        tr.step(collate([data[10 + step]]))
    torch.cuda.synchronize()
    res["ms_per_step"] = (time.time() - t0) / 5 * 1e3
    if rank == 0:
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"zero2_w{world}{tag}.json").write_text(json.dumps(res, indent=1))
        print(json.dumps({k: v for k, v in res.items() if k != "sha256"}))
    if world > 1:
        dist.destroy_process_group()


def compare(a, b):
    ra, rb = (json.loads((ROOT / "gpurun_out" / f"zero2_w{w}.json").read_text()) for w in (a, b))
    bad = [k for k in ra["sha256"] if ra["sha256"][k] != rb["sha256"][k]]
    print("store sha", ra.get("store_sha"), rb.get("store_sha"))
    for x, y in zip(ra.get("trace", []), rb.get("trace", [])):
        print({k: (x[k], y[k]) for k in x if x[k] != y[k]})
    for k in bad[:2]:
        print(k, "weights", ra["sums"][k], rb["sums"][k], "grads", ra["grad_sums"][k], rb["grad_sums"][k])
    print(json.dumps({"worlds": [a, b], "tensors": len(ra["sha256"]), "differing": bad, "loss": [ra["loss"], rb["loss"]],
                      "grad_norm": [ra["grad_norm"], rb["grad_norm"]], "state_fraction": [ra["optimizer_state_fraction"],
                                                                                        rb["optimizer_state_fraction"]]}))
    assert not bad, bad


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--compare":
        compare(sys.argv[2], sys.argv[3])
    else:
        main()
