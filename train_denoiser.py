"""`train_denoiser.py` — the reference's training entry point (train_denoiser.py:1621-1633: `python train_denoiser.py
cfg.yaml` under `accelerate launch` / torchrun) over the libb2f engine.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 train_denoiser.py scripts/denoiser/flux_qwen2p5vl_7b_vlm_stage2_512_synthetic.yaml

What runs per step (gpt_image_edit_b200/training.py: Stage2Trainer.step, reference :829-1181): VAE-encode target and
context, flow-matching noising with the resolution-shifted logit-normal sigma, frozen Qwen2.5-VL prefill, MLP2, FLUX
forward with block checkpoints, masked MSE against (noise - x0), backward with per-block recompute, ZeRO-2 gradient
reduce-scatter overlapped with the backward, global-norm clipping, AdamW on the rank's fp32 slice, bf16 all-gather.
One process per GPU over NCCL (RANK / LOCAL_RANK / WORLD_SIZE from the launcher), `seed + rank` data streams.

Replaces accelerate + DeepSpeed (scripts/accelerate_configs/zero2.json) and torch.autograd of the reference.
Out of scope, as SURVEY.md section 2 marks it: the real dataset code (univa/dataset), wandb, EMA, SigLIP/MLP3 branches.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import time
from pathlib import Path

import torch


def build_models(conf, device):
    """(model, vae, pipe, empty_pooled): the frozen stack and the trainable tower (reference :301-478, 795-805)."""
    from gpt_image_edit_b200.text_encoders import encode_prompt
    from univa.serve import cli

    mc = conf.model_config
    model, _, _ = cli.load_main_model_and_processor(mc.pretrained_lvlm_name_or_path, device, mc.synthetic, mc.small)
    pipe, tokenizers, text_encoders = cli.load_pipe(model.denoise_tower.denoiser, mc.pretrained_denoiser_name_or_path, device,
                                                    mc.synthetic, mc.small)
    if mc.pretrained_mlp2_path:
        sd = torch.load(mc.pretrained_mlp2_path, map_location="cpu")
        sd = {k.split("denoise_projector.")[-1]: v for k, v in sd.items() if "denoise_projector" in k}
        model.denoise_tower.denoise_projector.load_state_dict(sd)
    _, empty_pooled = encode_prompt(text_encoders, tokenizers, "", 256, device, 1)      # :795-805
    if conf.training_config.drop_t5_rate != 1.0:
        raise SystemExit("drop_t5_rate < 1 needs the T5 prompt embeddings of real captions; the stage yamls use 1.0")
    pipe.text_encoder = pipe.text_encoder_2 = None                                       # :806-808 (frees T5 / CLIP)
    del text_encoders
    torch.cuda.empty_cache()
    return model, pipe.vae, pipe, empty_pooled.to(torch.bfloat16)


def save_checkpoint(conf, trainer, model, step: int, rank: int, world: int):
    """checkpoint-{step}/: the trainable tensors under their diffusers names (safetensors, rank 0), denoise_projector.bin
    (reference :1231-1236) and this rank's optimizer partition (ZeRO: one file per rank, :1229 accelerator.save_state)."""
    from gpt_image_edit_b200 import checkpoint as ck

    tc = conf.training_config
    out = Path(tc.output_dir)
    out.mkdir(parents=True, exist_ok=True)
    if rank == 0 and tc.checkpoints_total_limit is not None:
        old = sorted((d for d in os.listdir(out) if d.startswith("checkpoint")), key=lambda x: int(x.split("-")[1]))
        if len(old) >= tc.checkpoints_total_limit:
            for d in old[:len(old) - tc.checkpoints_total_limit + 1]:
                shutil.rmtree(out / d)
    save = out / f"checkpoint-{step}"
    save.mkdir(parents=True, exist_ok=True)
    if rank == 0:
        den = model.denoise_tower.denoiser.state_dict()
        from gpt_image_edit_b200.training import check_param_is_in_components, get_trainable_params
        comps = get_trainable_params(conf.model_config.flux_train_layer_idx, model.denoise_tower.denoiser.config.num_layers,
                                     conf.model_config.only_tune_image_branch)
        trained = {k: v for k, v in den.items() if check_param_is_in_components("denoise_tower.denoiser." + k, comps)}
        ck.save_state_dict(trained, save / "denoiser_trainable")
        if conf.model_config.only_tune_mlp2 or conf.model_config.with_tune_mlp2:
            proj = {f"denoise_tower.denoise_projector.{k}": v.cpu() for k, v in
                    model.denoise_tower.denoise_projector.state_dict().items()}
            torch.save(proj, save / "denoise_projector.bin")
        (save / "trainer_state.json").write_text(json.dumps({"global_step": step, "world_size": world}))
    torch.save(trainer.opt.state_dict(), save / f"optimizer_rank{rank}.pt")
    return save


def main(conf):
    from gpt_image_edit_b200 import distributed as D
    from gpt_image_edit_b200.training import Stage2Trainer
    from univa.training.synthetic_data import SyntheticEditDataset, collate

    if not torch.cuda.is_available():
        raise SystemExit("train_denoiser.py runs on B200s through libb2f; there is no CPU path")
    tc, dc, mc = conf.training_config, conf.dataset_config, conf.model_config
    world, rank, local_rank = D.env_world()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    D.init_from_env(device=device)
    torch.manual_seed(tc.seed + rank)                                  # set_seed(seed, device_specific=True) (:290)
    if dc.dataset_type != "synthetic":
        raise SystemExit(f"dataset_type={dc.dataset_type!r}: only the synthetic triples of BASELINE.json configs[3] are "
                         "available offline (the reference's dataset code is out of scope)")
    model, vae, pipe, empty_pooled = build_models(conf, device)
    if world > 1:      # every rank starts from rank 0's weights (DeepSpeed broadcasts parameters at initialize())
        D.broadcast_weights(list(model.denoise_tower.denoiser._store.values()) +
                            list(model.denoise_tower.denoise_projector.state_dict().values()))
    trainer = Stage2Trainer(model, vae, pipe, tc, mc, empty_pooled)
    trainer.gen = torch.Generator(device=device).manual_seed(tc.seed + rank)
    n_train = sum(p.storage.numel() for p in trainer.params)
    if rank == 0:
        print(f"trainable tensors: {len(trainer.params)}  parameters: {n_train / 1e9:.3f} B  world: {world}  "
              f"ZeRO-2 buckets: {sum(b is not None for b in trainer.opt.buckets)}", flush=True)
    start = 0
    if tc.resume_from_checkpoint:
        ck = Path(tc.resume_from_checkpoint)
        trainer.opt.load_state_dict(torch.load(ck / f"optimizer_rank{rank}.pt", map_location=device))
        start = json.loads((ck / "trainer_state.json").read_text())["global_step"]
        trainer.global_step = start
    data = SyntheticEditDataset(dc.height, dc.width, dc.synthetic_len, seed=tc.seed + rank, target_sizes=dc.synthetic_target_sizes)
    loader = torch.utils.data.DataLoader(data, batch_size=dc.batch_size, collate_fn=collate, num_workers=0,
                                         pin_memory=dc.pin_memory)
    max_steps = tc.max_train_steps or (len(loader) * tc.num_train_epochs // tc.gradient_accumulation_steps)
    t0 = time.time()
    for batch in loader:
        out = trainer.step(batch)
        if not out["stepped"]:
            continue
        step = trainer.global_step
        loss = out["loss"].detach().clone()
        if world > 1:                                                   # accelerator.gather(loss) (:1168)
            torch.distributed.all_reduce(loss)
            loss /= world
        if rank == 0:
            dt = time.time() - t0
            print(f"step {step}  loss {loss.item():.5f}  grad_norm {out['grad_norm'].item():.4f}  lr {out['lr']:.3e}  "
                  f"{dc.batch_size * tc.gradient_accumulation_steps * world * (step - start) / max(dt, 1e-9):.3f} samples/s",
                  flush=True)
        if step % tc.checkpointing_steps == 0:
            path = save_checkpoint(conf, trainer, model, step, rank, world)
            if rank == 0:
                print(f"Saved state to {path}", flush=True)
        if step >= max_steps:
            break
    D.barrier()
    return trainer


if __name__ == "__main__":
    from univa.training.configuration_denoise import load_config

    parser = argparse.ArgumentParser()
    parser.add_argument("config", type=str)
    main(load_config(parser.parse_args().config))
