"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/flux_oracle.py for the rules).

Restatement of the reference sampling loop (univa/utils/flux_pipeline.py:874-1130, source on disk)
and of diffusers 0.32.2 `FlowMatchEulerDiscreteScheduler` in plain torch, device- and
dtype-agnostic, wired to the oracle transformer/VAE.  This is (a) the checker for the product
pipeline and (b) the CPU baseline `bench.py --impl reference` times (the reference's own diffusers
path cannot be imported here: diffusers is absent).

Pin: `sample()` is bit-identical (fp32) to the reference's own `FluxKontextPipeline.__call__` executed from
/root/reference/univa/utils/flux_pipeline.py with stand-ins for the diffusers base classes
(tests/golden/make_pipeline_ref_golden.py -> tests/golden/pipeline_ref_loop.pt, checked by
tests/test_oracle_cpu.py).  The scheduler arithmetic itself (diffusers, not on disk) remains unpinned by the reference.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import flux_oracle as fo
from . import vae_oracle as vo


# ----------------------------------------------------------------------------- scheduler (A.5)
class EulerSchedulerOracle:
    order = 1

    def __init__(self, num_train_timesteps=1000, base_shift=0.5, max_shift=1.15, base_image_seq_len=256,
                 max_image_seq_len=4096):
        self.config = dict(num_train_timesteps=num_train_timesteps, base_shift=base_shift, max_shift=max_shift,
                           base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len,
                           use_dynamic_shifting=True)
        self._step_index = None
        self._begin_index = None

    def set_timesteps(self, sigmas, mu, device=None):
        sigmas = np.array(sigmas).astype(np.float32)
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)      # time_shift(mu, 1.0, sigmas)
        s = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = s * self.config["num_train_timesteps"]
        self.sigmas = torch.cat([s, torch.zeros(1, device=s.device)])
        self._step_index = None
        self._begin_index = None

    def set_begin_index(self, i=0):
        self._begin_index = i

    def step(self, model_output, timestep, sample):
        if self._step_index is None:
            self._step_index = self._begin_index if self._begin_index is not None else \
                int((self.timesteps == timestep).nonzero()[0])
        x = sample.to(torch.float32)
        sigma, sigma_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = x + (sigma_next - sigma) * model_output   # 0-dim fp32 * tensor: evaluated in model_output.dtype
        self._step_index += 1
        return prev.to(model_output.dtype)


def calculate_shift(n, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)   # reference flux_pipeline.py:106-116
    b = base_shift - m * base_seq_len
    return n * m + b


# ----------------------------------------------------------------------------- layout (reference :561-598)
def latent_image_ids(h, w):
    ids = torch.zeros(h, w, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w)[None, :]
    return ids.reshape(h * w, 3)


def pack_latents(x):
    B, C, H, W = x.shape
    return x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)


def unpack_latents(x, h_lat, w_lat):
    B, _, ch = x.shape
    return x.view(B, h_lat // 2, w_lat // 2, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, ch // 4, h_lat, w_lat)


def target_size(height, width, max_area, mult=16):
    aspect = width / height                                      # reference :874-884
    width = round((max_area * aspect) ** 0.5)
    height = round((max_area / aspect) ** 0.5)
    return height // mult * mult, width // mult * mult


# ----------------------------------------------------------------------------- the loop (reference :976-1130)
@torch.no_grad()
def sample(flux_sd, flux_cfg: fo.FluxConfig, vae_sd, vae_cfg: vo.VaeConfig, image, prompt_embeds, pooled, *,
           height, width, num_inference_steps=28, guidance_scale=3.5, latents=None, max_area=None, output="image",
           callback=None, true_cfg_scale=1.0, negative_prompt_embeds=None, negative_pooled=None):
    """image: [B,3,H,W] in [-1,1] (already at its final resolution, i.e. `_auto_resize=False`);
    latents: packed initial noise [B,S_tgt,64] (the run is deterministic given it).
    Returns decoded image [B,3,H,W] (output="image") or the final packed latents (output="latent")."""
    dtype, device = prompt_embeds.dtype, prompt_embeds.device
    height, width = target_size(height, width, max_area if max_area is not None else height * width)
    B = prompt_embeds.shape[0]
    h_lat, w_lat = 2 * (height // 16), 2 * (width // 16)
    z_img = vo.encode_mode(vae_sd, vae_cfg, image.to(dtype))
    z_img = (z_img - vae_cfg.shift_factor) * vae_cfg.scaling_factor
    image_latents = pack_latents(z_img)
    image_ids = latent_image_ids(z_img.shape[2] // 2, z_img.shape[3] // 2)
    image_ids[..., 0] = 1
    latent_ids = torch.cat([latent_image_ids(h_lat // 2, w_lat // 2), image_ids], dim=0).to(device=device, dtype=dtype)
    text_ids = torch.zeros(prompt_embeds.shape[1], 3, device=device, dtype=dtype)
    latents = latents.to(device=device, dtype=dtype)

    sched = EulerSchedulerOracle()
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    sched.set_timesteps(sig, calculate_shift(latents.shape[1]), device=device)
    guidance = torch.full([1], guidance_scale, device=device, dtype=torch.float32).expand(B) \
        if flux_cfg.guidance_embeds else None
    sched.set_begin_index(0)
    for i, t in enumerate(sched.timesteps):
        x_in = torch.cat([latents, image_latents], dim=1)
        timestep = t.expand(B).to(latents.dtype)
        v = fo.flux_forward(flux_sd, flux_cfg, x_in, prompt_embeds, pooled, timestep / 1000, latent_ids, text_ids,
                            guidance=guidance)
        v = v[:, : latents.size(1)]
        if true_cfg_scale > 1 and negative_prompt_embeds is not None:          # reference flux_pipeline.py:1080-1095
            neg_ids = torch.zeros(negative_prompt_embeds.shape[1], 3, device=device, dtype=dtype)
            vn = fo.flux_forward(flux_sd, flux_cfg, x_in, negative_prompt_embeds, negative_pooled, timestep / 1000, latent_ids,
                                 neg_ids, guidance=guidance)[:, : latents.size(1)]
            v = vn + true_cfg_scale * (v - vn)
        latents = sched.step(v, t, latents)
        if callback is not None:
            callback(i, latents)
    if output == "latent":
        return latents
    z = unpack_latents(latents, h_lat, w_lat)
    z = (z / vae_cfg.scaling_factor) + vae_cfg.shift_factor
    return vo.decode(vae_sd, vae_cfg, z.to(dtype))
