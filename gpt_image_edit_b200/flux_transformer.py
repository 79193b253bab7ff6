"""Drop-in for the object the reference passes as `transformer=` to FluxKontextPipeline
(`model.denoise_tower.denoiser`, a diffusers `FluxTransformer2DModel`; reference
univa/serve/cli.py:64-68, univa/models/modeling_univa_denoise_tower.py:21) — same call signature,
config attributes and state-dict key names, but the forward is a single C-ABI call into libb2f
(`b2f_flux_forward`, hand-written sm_100a kernels).  Nothing here computes the model in torch.

Storage: projections that the engine runs as one GEMM (q/k/v, the single block's q/k/v/proj_mlp,
every AdaLN linear) are STORED row-concatenated; `state_dict()` / `load_state_dict()` expose and
accept the diffusers names as views, so checkpoints interchange.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from types import SimpleNamespace

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class FluxTransformerConfig(SimpleNamespace):
    """Attribute names follow diffusers FluxTransformer2DModel.config (the pipeline reads
    `.in_channels` and `.guidance_embeds`: reference flux_pipeline.py:975, :1011)."""

    def __init__(self, **kw):
        base = dict(patch_size=1, in_channels=64, out_channels=64, num_layers=19, num_single_layers=38,
                    attention_head_dim=128, num_attention_heads=24, joint_attention_dim=4096,
                    pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
        base.update(kw)
        super().__init__(**base)

    def get(self, k, default=None):
        return getattr(self, k, default)


class Transformer2DModelOutput(SimpleNamespace):
    pass


def _fused_layout(cfg: FluxTransformerConfig):
    """fused name -> list of (diffusers linear name, out_features); plus plain names."""
    d = cfg.num_attention_heads * cfg.attention_head_dim
    fused: "OrderedDict[str, list]" = OrderedDict()
    adaln = []
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        adaln += [(p + "norm1.linear", 6 * d), (p + "norm1_context.linear", 6 * d)]
        fused[p + "attn.qkv"] = [(p + "attn.to_q", d), (p + "attn.to_k", d), (p + "attn.to_v", d)]
        fused[p + "attn.add_qkv"] = [(p + "attn.add_q_proj", d), (p + "attn.add_k_proj", d), (p + "attn.add_v_proj", d)]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        adaln += [(p + "norm.linear", 3 * d)]
        fused[p + "qkv_mlp"] = [(p + "attn.to_q", d), (p + "attn.to_k", d), (p + "attn.to_v", d), (p + "proj_mlp", 4 * d)]
    adaln += [("norm_out.linear", 2 * d)]
    fused["adaln"] = adaln
    return fused


def _plain_linears(cfg: FluxTransformerConfig):
    d = cfg.num_attention_heads * cfg.attention_head_dim
    out = [("x_embedder", d, cfg.in_channels), ("context_embedder", d, cfg.joint_attention_dim),
           ("time_text_embed.timestep_embedder.linear_1", d, 256), ("time_text_embed.timestep_embedder.linear_2", d, d)]
    if cfg.guidance_embeds:
        out += [("time_text_embed.guidance_embedder.linear_1", d, 256), ("time_text_embed.guidance_embedder.linear_2", d, d)]
    out += [("time_text_embed.text_embedder.linear_1", d, cfg.pooled_projection_dim),
            ("time_text_embed.text_embedder.linear_2", d, d), ("proj_out", cfg.out_channels, d)]
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        out += [(p + "attn.to_out.0", d, d), (p + "attn.to_add_out", d, d), (p + "ff.net.0.proj", 4 * d, d),
                (p + "ff.net.2", d, 4 * d), (p + "ff_context.net.0.proj", 4 * d, d), (p + "ff_context.net.2", d, 4 * d)]
    for i in range(cfg.num_single_layers):
        out += [(f"single_transformer_blocks.{i}.proj_out", d, 5 * d)]
    return out


def _norm_weights(cfg: FluxTransformerConfig):
    out = []
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}.attn."
        out += [p + "norm_q.weight", p + "norm_k.weight", p + "norm_added_q.weight", p + "norm_added_k.weight"]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}.attn."
        out += [p + "norm_q.weight", p + "norm_k.weight"]
    return out


class B200FluxTransformer2DModel(torch.nn.Module):
    def __init__(self, config: FluxTransformerConfig | None = None, device="cuda", **kw):
        super().__init__()
        self.config = config or FluxTransformerConfig(**kw)
        cfg = self.config
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.B2FError("B200FluxTransformer2DModel lives on a CUDA device; there is no CPU path")
        self.inner_dim = d = cfg.num_attention_heads * cfg.attention_head_dim
        self._store: "OrderedDict[str, torch.Tensor]" = OrderedDict()   # bound name -> storage tensor
        self._views: "OrderedDict[str, torch.Tensor]" = OrderedDict()   # diffusers name -> view
        mk = lambda *shape: torch.zeros(shape, device=dev, dtype=torch.bfloat16)
        for fname, parts in _fused_layout(cfg).items():
            rows = sum(n for _, n in parts)
            w, b = mk(rows, d), mk(rows)
            self._store[fname + ".weight"], self._store[fname + ".bias"] = w, b
            r = 0
            for name, n in parts:
                self._views[name + ".weight"], self._views[name + ".bias"] = w[r:r + n], b[r:r + n]
                r += n
        for name, o, i in _plain_linears(cfg):
            w, b = mk(o, i), mk(o)
            self._store[name + ".weight"], self._store[name + ".bias"] = w, b
            self._views[name + ".weight"], self._views[name + ".bias"] = w, b
        for name in _norm_weights(cfg):
            w = torch.ones(cfg.attention_head_dim, device=dev, dtype=torch.bfloat16)
            self._store[name] = w
            self._views[name] = w
        for k, t in self._store.items():   # registered so .parameters()/.to() bookkeeping sees them
            self.register_buffer("w__" + k.replace(".", "__"), t, persistent=False)

        ccfg = _lib.FluxCfg(cfg.num_attention_heads, cfg.attention_head_dim, cfg.num_layers, cfg.num_single_layers,
                            cfg.in_channels, cfg.out_channels, cfg.joint_attention_dim, cfg.pooled_projection_dim,
                            int(bool(cfg.guidance_embeds)), 4)
        h = C.c_void_p()
        check(_lib.lib.b2f_flux_create(C.byref(h), C.byref(ccfg)), "b2f_flux_create")
        self._h = h
        for k, t in self._store.items():
            check(_lib.lib.b2f_flux_bind_weight(self._h, k.encode(), ptr(t), t.numel()), f"bind {k}")
        check(_lib.lib.b2f_flux_finalize(self._h), "b2f_flux_finalize")
        self.mod_width = int(_lib.lib.b2f_flux_mod_width(self._h))
        self._ws = {}
        self._rope = None
        self._schedule = None
        self.gradient_checkpointing = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None and getattr(_lib, "lib", None) is not None:  # not during interpreter teardown
            _lib.lib.b2f_flux_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ nn.Module surface
    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        return next(iter(self._store.values())).device

    def state_dict(self, *a, **k):
        return OrderedDict((n, t) for n, t in self._views.items())

    def load_state_dict(self, sd, strict: bool = True, assign: bool = False):
        missing = [k for k in self._views if k not in sd]
        unexpected = [k for k in sd if k not in self._views]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]}... unexpected {unexpected[:5]}...")
        with torch.no_grad():
            for k, v in self._views.items():
                if k in sd:
                    if tuple(sd[k].shape) != tuple(v.shape):
                        raise RuntimeError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(v.shape)}")
                    v.copy_(sd[k])
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    @torch.no_grad()
    def randomize_(self, seed: int = 0, std: float = 0.02, bias_std: float = 0.02):
        """Seeded synthetic weights drawn on the device (no checkpoints exist offline):
        weights ~ N(0, std^2), biases ~ N(0, bias_std^2), RMSNorm weights = 1."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        for k, t in self._store.items():
            if k.endswith("norm_q.weight") or k.endswith("norm_k.weight") or "norm_added" in k:
                t.fill_(1.0)
            else:
                s_ = bias_std if k.endswith(".bias") else std
                # draw in fp32 chunks to bound temporary memory for the multi-GB fused tensors
                flat = t.view(-1)
                step = 1 << 26
                for o in range(0, flat.numel(), step):
                    n = min(step, flat.numel() - o)
                    flat[o:o + n] = (torch.randn(n, device=self.device, generator=g) * s_).to(torch.bfloat16)
        return self

    def named_parameters_diffusers(self):
        yield from self._views.items()

    def enable_gradient_checkpointing(self):  # reference train_denoiser.py:486
        self.gradient_checkpointing = True

    def to(self, *args, **kwargs):  # the pipeline calls pipe.to(device) (reference cli.py:69)
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a).type != "cuda":
                raise _lib.B2FError("B200FluxTransformer2DModel cannot leave the GPU: there is no CPU path")
            if isinstance(a, torch.dtype) and a != torch.bfloat16:
                raise _lib.B2FError("B200FluxTransformer2DModel computes in bfloat16 only")
        return self

    # ------------------------------------------------------------------ helpers
    def _workspace(self, nbytes: int, tag) -> torch.Tensor:
        buf = self._ws.get(tag)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[tag] = buf
        return buf

    def _temb_mod(self, t1000: torch.Tensor, g1000: torch.Tensor | None, pooled: torch.Tensor, want_silu: bool = False):
        """rows of (timestep*1000, guidance*1000) fp32 + pooled bf16 -> (temb, mod) via the C ABI
        (want_silu: also silu(temb), the input of every AdaLN linear — the training step needs it)."""
        rows = t1000.numel()
        d = self.inner_dim
        temb = torch.empty((rows, d), device=self.device, dtype=torch.bfloat16)
        stemb = torch.empty_like(temb)
        mod = torch.empty((rows, self.mod_width), device=self.device, dtype=torch.bfloat16)
        nws = int(_lib.lib.b2f_flux_temb_workspace_bytes(self._h, rows))
        ws = self._workspace(nws, "temb")
        check(_lib.lib.b2f_flux_temb(self._h, ptr(t1000), ptr(g1000), ptr(pooled), pooled.stride(0), rows,
                                     ptr(temb), ptr(stemb), ptr(ws), nws, stream_ptr()), "b2f_flux_temb")
        check(_lib.lib.b2f_flux_modulation(self._h, ptr(stemb), rows, ptr(mod), stream_ptr()), "b2f_flux_modulation")
        if want_silu:
            return temb, mod, stemb
        return temb, mod

    @staticmethod
    def _times1000(x: torch.Tensor) -> torch.Tensor:
        # diffusers: `timestep.to(hidden_states.dtype) * 1000` in bf16, then the
        # sinusoid uses `.float()` of that value.  Scalar bookkeeping, kept in torch on purpose.
        return (x.to(torch.bfloat16) * 1000).float().contiguous()

    def prepare_schedule(self, timesteps_over_1000: torch.Tensor, guidance: torch.Tensor | None,
                         pooled_projections: torch.Tensor):
        """Hoist the AdaLN modulation of a whole sampling schedule (one weight-streaming GEMM
        with M = steps*B instead of `steps` GEMMs with M = B).  timesteps_over_1000: [n_steps]
        values exactly as the loop will pass them (bf16, already divided by 1000)."""
        n = timesteps_over_1000.numel()
        B = pooled_projections.shape[0]
        t = self._times1000(timesteps_over_1000.reshape(n, 1).expand(n, B).reshape(-1))
        g = None
        if self.config.guidance_embeds:
            g = self._times1000(guidance.reshape(1, B).expand(n, B).reshape(-1))
        pooled = pooled_projections.to(torch.bfloat16).repeat(n, 1).contiguous()
        _, mod = self._temb_mod(t, g, pooled)
        self._schedule = SimpleNamespace(n=n, B=B, mod=mod.view(n, B, self.mod_width))
        return self._schedule

    def _set_rope(self, txt_ids, img_ids, S_txt: int, S_img: int):
        """FluxPosEmbed tables for [txt; img] ids, handed to the engine (rebuilt only when the ids change)."""
        if txt_ids.dim() == 3:
            txt_ids = txt_ids[0]
        if img_ids.dim() == 3:
            img_ids = img_ids[0]
        key = (txt_ids.data_ptr(), img_ids.data_ptr(), S_txt, S_img, txt_ids._version, img_ids._version)
        if self._rope is None or self._rope[0] != key:
            from . import ops
            ids = torch.cat((txt_ids.float(), img_ids.float()), dim=0).contiguous()
            cos, sin = ops.rope_tables(ids, tuple(self.config.axes_dims_rope))
            self._rope = (key, cos, sin, (txt_ids, img_ids))
        _, cos, sin, _keep = self._rope
        check(_lib.lib.b2f_flux_set_rope(self._h, ptr(cos), ptr(sin), S_txt + S_img), "b2f_flux_set_rope")

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None,
                img_ids=None, txt_ids=None, guidance=None, joint_attention_kwargs=None, return_dict=True,
                **unused):
        cfg = self.config
        jak = dict(joint_attention_kwargs or {})
        if jak.get("attention_mask") is not None:
            raise _lib.B2FError("attention_mask (mixed-size training batches) is not implemented in libb2f")
        B, S_img, _ = hidden_states.shape
        S_txt = encoder_hidden_states.shape[1]
        S = S_txt + S_img
        hs = hidden_states.to(torch.bfloat16).contiguous()
        enc = encoder_hidden_states.to(torch.bfloat16).contiguous()
        self._set_rope(txt_ids, img_ids, S_txt, S_img)

        step = jak.get("_b2f_schedule_step")
        if step is not None and self._schedule is not None and self._schedule.B == B:
            mod = self._schedule.mod[int(step)]
        else:
            t = self._times1000(timestep.reshape(-1).expand(B))
            g = self._times1000(guidance.reshape(-1).expand(B)) if cfg.guidance_embeds else None
            pooled = pooled_projections.to(torch.bfloat16).contiguous()
            _, mod = self._temb_mod(t, g, pooled)
        n_out = int(jak.get("_b2f_out_rows", S_img))
        out = torch.empty((B, n_out, cfg.out_channels), device=self.device, dtype=torch.bfloat16)
        nws = int(_lib.lib.b2f_flux_workspace_bytes(self._h, B, S_img, S_txt))
        ws = self._workspace(nws, "fwd")
        first, last = jak.get("_b2f_block_range", (0, -1))
        check(_lib.lib.b2f_flux_forward(self._h, ptr(hs), ptr(enc), ptr(mod), mod.stride(0), ptr(out), B, S_img,
                                        S_txt, n_out, ptr(ws), nws, int(first), int(last), stream_ptr()),
              "b2f_flux_forward")
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)

    def debug_hidden(self, B: int, S_img: int, S_txt: int) -> torch.Tensor:
        """View of the joint activation buffer h[B, S_txt+S_img, d] inside the workspace (block-level
        parity tests read it after a partial-range forward)."""
        ws = self._ws["fwd"]
        off = (-ws.data_ptr()) % 256
        n = B * (S_img + S_txt) * self.inner_dim
        return ws[off:off + 2 * n].view(torch.bfloat16).view(B, S_img + S_txt, self.inner_dim)
