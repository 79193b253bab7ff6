"""FlowMatchEulerDiscreteScheduler for the FLUX-Kontext loop — host-side schedule arithmetic plus
the Euler update through libb2f (`b2f_euler_step`).

Drop-in for the object the reference pipeline drives (reference univa/utils/flux_pipeline.py:995-1007
`retrieve_timesteps(scheduler, sigmas=, mu=)`, :1052 `set_begin_index(0)`, :1099 `step(v, t, x)`);
the arithmetic restates diffusers 0.32.2.  The integer step index is host state and
advances exactly as in diffusers (bit-exact by construction).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from ._lib import B2FError


class _Config(dict):
    __getattr__ = dict.get


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 3.0, use_dynamic_shifting: bool = True,
                 base_shift: float = 0.5, max_shift: float = 1.15, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 4096):
        self.config = _Config(num_train_timesteps=num_train_timesteps, shift=shift,
                              use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift, max_shift=max_shift,
                              base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len)
        self.timesteps = None
        self.sigmas = None
        self._sigmas_host = None
        self._step_index = None
        self._begin_index = None
        self.num_inference_steps = None

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    @staticmethod
    def time_shift(mu: float, sigma: float, t: np.ndarray) -> np.ndarray:
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        c = self.config
        if c.use_dynamic_shifting and mu is None:
            raise ValueError("you have to pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            ts = np.linspace(c.num_train_timesteps, 1.0, num_inference_steps)
            sigmas = ts / c.num_train_timesteps
        else:
            sigmas = np.array(sigmas).astype(np.float32)
            num_inference_steps = len(sigmas)
        self.num_inference_steps = num_inference_steps
        if c.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = c.shift * sigmas / (1 + (c.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        self._sigmas_host = torch.cat([sig, torch.zeros(1)])           # host copy: dt without a device sync
        self.timesteps = (sig * c.num_train_timesteps).to(device=device)
        self.sigmas = self._sigmas_host.to(device=device)
        self._step_index = None
        self._begin_index = None

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            t = float(timestep)
            idx = (self.timesteps.cpu() == t).nonzero()
            self._step_index = int(idx[1 if len(idx) > 1 else 0])
        else:
            self._step_index = self._begin_index

    def dt(self, i: int) -> float:
        """sigma[i+1] - sigma[i] evaluated in fp32, as the 0-dim fp32 tensor subtraction in diffusers."""
        s = self._sigmas_host
        return float((s[i + 1] - s[i]).item())

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, **kw):
        """x <- bf16(float(x) + bf16(bf16(dt)*v)); in place on `sample` (returned).  CUDA bf16 tensors only."""
        if self._step_index is None:
            self._init_step_index(timestep)
        if not (sample.is_cuda and sample.dtype == torch.bfloat16 and model_output.dtype == torch.bfloat16):
            raise B2FError("FlowMatchEulerDiscreteScheduler.step runs through libb2f: CUDA bf16 tensors only")
        ops.euler_step_(sample, model_output, self.dt(self._step_index))
        self._step_index += 1
        return (sample,) if not return_dict else _Config(prev_sample=sample)
