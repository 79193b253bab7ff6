"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (gpt_image_edit_b200/, univa/, train_denoiser.py).

Plain-torch restatements of the optimiser arithmetic the reference reaches through accelerate + DeepSpeed + torch
(train_denoiser.py:596-602 AdamW, :1174-1177 clip_grad_norm_, scripts/accelerate_configs/zero2.json):
  * `TorchMath`: drop-in for `training._B2FMath` so that the CPU `gloo` tests can drive `ShardedAdamW`'s partitioning
    logic without a GPU (the product always uses the CUDA kernels);
  * `reference_step`: the un-partitioned update — average the ranks' gradients, clip by the global norm, torch AdamW
    on fp32 master weights, round to bf16 — what ZeRO-2 must reproduce on every rank.
"""
from __future__ import annotations

import torch


class TorchMath:
    @staticmethod
    def cast_to_f32(src):
        return src.float().clone()

    @staticmethod
    def cast_to_bf16(src):
        return src.to(torch.bfloat16)

    @staticmethod
    def sumsq(g, out, accumulate):
        s = g.double().pow(2).sum().float()
        out.copy_(out + s if accumulate else s.reshape(1))

    @staticmethod
    def clip_coef(sumsq, max_norm, pre_scale):
        norm = sumsq.sqrt() * pre_scale
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0) * pre_scale if max_norm > 0 else torch.full_like(norm, pre_scale)
        return coef, norm

    @staticmethod
    def adamw(p32, m, v, g, p16, *, lr, betas, eps, weight_decay, step, gscale):
        g = g * gscale
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        p32.mul_(1 - lr * weight_decay)
        p32.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + eps, value=-lr / bc1)
        if p16 is not None:
            p16.copy_(p32.to(torch.bfloat16))


def reference_step(weights_bf16, grads_per_rank, steps_state=None, *, lr, betas, eps, weight_decay, max_grad_norm):
    """weights_bf16: list of bf16 tensors; grads_per_rank: [rank][param] fp32.  Returns (new bf16 weights, grad norm,
    state) using torch.optim.AdamW on fp32 masters and torch.nn.utils.clip_grad_norm_."""
    if steps_state is None:
        masters = [torch.nn.Parameter(w.float().clone()) for w in weights_bf16]
        opt = torch.optim.AdamW(masters, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        steps_state = (masters, opt)
    masters, opt = steps_state
    world = len(grads_per_rank)
    for i, p in enumerate(masters):
        p.grad = sum(g[i] for g in grads_per_rank) / world
    norm = torch.nn.utils.clip_grad_norm_(masters, max_grad_norm)
    for g in opt.param_groups:
        g["lr"] = lr
    opt.step()
    return [p.data.to(torch.bfloat16) for p in masters], norm, steps_state
