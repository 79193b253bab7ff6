"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/flux_oracle.py for the rules).

PyTorch restatement of diffusers==0.32.2 `AutoencoderKL` with the FLUX VAE config:
block_out_channels (128,256,512,512), layers_per_block 2, latent_channels 16, GroupNorm(32, eps 1e-6),
SiLU, mid-block single-head attention, no quant convs.  State-dict keys are the diffusers names.
Reference call sites: univa/utils/flux_pipeline.py:600-613 (encode → mode → affine), :1127-1129
(decode), train_denoiser.py:887-898.

Parity pin: cross-checked against torchtitan.experiments.flux.model.autoencoder (independent
BFL-layout implementation) through the key mapping in tests/golden/make_golden.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    block_out_channels: tuple = (128, 256, 512, 512)
    layers_per_block: int = 2
    latent_channels: int = 16
    norm_num_groups: int = 32
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159

    def get(self, k, default=None):
        return getattr(self, k, default)


def state_dict_spec(cfg: VaeConfig) -> dict[str, tuple]:
    s: dict[str, tuple] = {}

    def conv(name, o, i, k):
        s[name + ".weight"] = (o, i, k, k)
        s[name + ".bias"] = (o,)

    def norm(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)

    def resnet(name, i, o):
        norm(name + ".norm1", i)
        conv(name + ".conv1", o, i, 3)
        norm(name + ".norm2", o)
        conv(name + ".conv2", o, o, 3)
        if i != o:
            conv(name + ".conv_shortcut", o, i, 1)

    def mid(name, c):
        resnet(name + ".resnets.0", c, c)
        norm(name + ".attentions.0.group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[f"{name}.attentions.0.{n}.weight"] = (c, c)
            s[f"{name}.attentions.0.{n}.bias"] = (c,)
        resnet(name + ".resnets.1", c, c)

    boc = cfg.block_out_channels
    conv("encoder.conv_in", boc[0], cfg.in_channels, 3)
    ch = boc[0]
    for i, o in enumerate(boc):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", ch, o)
            ch = o
        if i != len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", o, o, 3)
    mid("encoder.mid_block", boc[-1])
    norm("encoder.conv_norm_out", boc[-1])
    conv("encoder.conv_out", 2 * cfg.latent_channels, boc[-1], 3)

    rev = list(reversed(boc))
    conv("decoder.conv_in", rev[0], cfg.latent_channels, 3)
    mid("decoder.mid_block", rev[0])
    ch = rev[0]
    for i, o in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", ch, o)
            ch = o
        if i != len(rev) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", o, o, 3)
    norm("decoder.conv_norm_out", rev[-1])
    conv("decoder.conv_out", cfg.out_channels, rev[-1], 3)
    return s


def make_synthetic_state_dict(cfg: VaeConfig, seed: int = 0, dtype=torch.float32, device="cpu"):
    """Kaiming-scaled random weights (keeps activations O(1) through ~30 conv layers), random norm
    affine and biases so that dropping any of them is visible."""
    sd = {}
    for idx, (name, shape) in enumerate(state_dict_spec(cfg).items()):
        g = torch.Generator().manual_seed(seed * 7_000_003 + idx)
        if name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            t = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        sd[name] = t.to(dtype=dtype, device=device)
    return sd


def _gn(sd, name, x, groups):
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps=1e-6)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, name, x, groups):
    """`ResnetBlock2D` (temb=None, output_scale_factor=1)."""
    h = _conv(sd, name + ".conv1", F.silu(_gn(sd, name + ".norm1", x, groups)))
    h = _conv(sd, name + ".conv2", F.silu(_gn(sd, name + ".norm2", h, groups)))
    if name + ".conv_shortcut.weight" in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return x + h


def mid_attention(sd, name, x, groups):
    """diffusers `Attention(heads=1, dim_head=C, residual_connection=True, norm_num_groups=32)` over H*W tokens."""
    B, C, H, W = x.shape
    res = x
    h = x.view(B, C, H * W)
    h = F.group_norm(h, groups, sd[name + ".group_norm.weight"], sd[name + ".group_norm.bias"], eps=1e-6)
    h = h.transpose(1, 2)
    lin = lambda n, t: F.linear(t, sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"])
    q, k, v = lin("to_q", h)[:, None], lin("to_k", h)[:, None], lin("to_v", h)[:, None]
    if x.is_cuda and x.dtype != torch.float64:
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        s = (q @ k.transpose(-1, -2)) * (C ** -0.5)
        o = torch.softmax(s, dim=-1) @ v
    o = lin("to_out.0", o[:, 0])
    return o.transpose(1, 2).reshape(B, C, H, W) + res


def _mid_block(sd, name, x, groups):
    x = resnet_block(sd, name + ".resnets.0", x, groups)
    x = mid_attention(sd, name + ".attentions.0", x, groups)
    return resnet_block(sd, name + ".resnets.1", x, groups)


def encode_moments(sd, cfg: VaeConfig, x):
    """`AutoencoderKL.encode(x)` up to the DiagonalGaussianDistribution parameters -> (mean, logvar)."""
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = _conv(sd, "encoder.conv_in", x)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
        if i != len(boc) - 1:
            h = F.pad(h, (0, 1, 0, 1))  # Downsample2D(padding=0): pad right/bottom then stride-2 conv
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = _mid_block(sd, "encoder.mid_block", h, g)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, g)))
    mean, logvar = torch.chunk(h, 2, dim=1)
    return mean, torch.clamp(logvar, -30.0, 20.0)


def encode_mode(sd, cfg, x):
    """`.latent_dist.mode()` — what inference uses (reference flux_pipeline.py:609, sample_mode="argmax")."""
    return encode_moments(sd, cfg, x)[0]


def encode_sample(sd, cfg, x, generator=None):
    """`.latent_dist.sample()` — what training uses (train_denoiser.py:887, 895)."""
    mean, logvar = encode_moments(sd, cfg, x)
    noise = torch.randn(mean.shape, generator=generator, device=mean.device, dtype=mean.dtype)
    return mean + torch.exp(0.5 * logvar) * noise


def decode(sd, cfg: VaeConfig, z):
    """`AutoencoderKL.decode(z, return_dict=False)[0]`."""
    g = cfg.norm_num_groups
    rev = list(reversed(cfg.block_out_channels))
    h = _conv(sd, "decoder.conv_in", z)
    h = _mid_block(sd, "decoder.mid_block", h, g)
    for i in range(len(rev)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
        if i != len(rev) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    return _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, g)))
