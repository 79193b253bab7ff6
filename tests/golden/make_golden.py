"""Generates tests/golden/*.pt — run HERE (build container), commit the outputs.

Pins for the oracle (the reference ships no golden vectors):
  flux_toy_titan.pt   inputs + output of torchtitan.experiments.flux.FluxModel (an independent
                      BFL-layout FLUX implementation that is importable in this image) in float64,
                      with the oracle's synthetic diffusers-named weights mapped by SURVEY.md A.7.
                      guidance_embeds=False because torchtitan has no guidance embedder.
  flux_toy_oracle.pt  inputs + fp32 output of the oracle itself WITH guidance (regression pin for
                      the pieces torchtitan cannot cover: guidance MLP, bf16 timestep chain).
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import flux_oracle as fo  # noqa: E402

HERE = Path(__file__).resolve().parent
TOY = dict(num_layers=2, num_single_layers=2, attention_head_dim=128, num_attention_heads=2,
           joint_attention_dim=256, pooled_projection_dim=64)


def toy_inputs(cfg, B=1, S_txt=24, HL=6, WL=5, seed=1, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    n = HL * WL
    ids = torch.zeros(HL, WL, 3)
    ids[..., 1] += torch.arange(HL)[:, None]
    ids[..., 2] += torch.arange(WL)[None, :]
    ids = ids.reshape(-1, 3)
    ctx = ids.clone()
    ctx[:, 0] = 1
    return dict(
        hidden_states=torch.randn(B, 2 * n, cfg.in_channels, generator=g, dtype=dtype),
        encoder_hidden_states=torch.randn(B, S_txt, cfg.joint_attention_dim, generator=g, dtype=dtype),
        pooled_projections=torch.randn(B, cfg.pooled_projection_dim, generator=g, dtype=dtype),
        timestep=torch.full((B,), 0.75, dtype=dtype),
        guidance=torch.full((B,), 3.5, dtype=dtype),
        img_ids=torch.cat([ids, ctx]),
        txt_ids=torch.zeros(S_txt, 3),
    )


def titan_from_diffusers(cfg: fo.FluxConfig, sd: dict):
    """Build torchtitan's FluxModel and load the diffusers-named weights through the A.7 mapping."""
    from torchtitan.experiments.flux.model.args import FluxModelArgs
    from torchtitan.experiments.flux.model.model import FluxModel

    d = cfg.inner_dim
    args = FluxModelArgs(in_channels=cfg.in_channels, out_channels=cfg.out_channels, vec_in_dim=cfg.pooled_projection_dim,
                         context_in_dim=cfg.joint_attention_dim, hidden_size=d, num_heads=cfg.num_attention_heads,
                         depth=cfg.num_layers, depth_single_blocks=cfg.num_single_layers, axes_dim=tuple(cfg.axes_dims_rope))
    m = FluxModel(args).double()
    cat = lambda *names: torch.cat([sd[n] for n in names], dim=0)
    t = {}

    def lin(dst, src):
        t[dst + ".weight"], t[dst + ".bias"] = sd[src + ".weight"], sd[src + ".bias"]

    lin("img_in", "x_embedder")
    lin("txt_in", "context_embedder")
    lin("time_in.in_layer", "time_text_embed.timestep_embedder.linear_1")
    lin("time_in.out_layer", "time_text_embed.timestep_embedder.linear_2")
    lin("vector_in.in_layer", "time_text_embed.text_embedder.linear_1")
    lin("vector_in.out_layer", "time_text_embed.text_embedder.linear_2")
    for i in range(cfg.num_layers):
        s, p = f"transformer_blocks.{i}.", f"double_blocks.{i}."
        lin(p + "img_mod.lin", s + "norm1.linear")
        lin(p + "txt_mod.lin", s + "norm1_context.linear")
        for kind in ("weight", "bias"):
            t[p + f"img_attn.qkv.{kind}"] = cat(*(s + f"attn.{n}.{kind}" for n in ("to_q", "to_k", "to_v")))
            t[p + f"txt_attn.qkv.{kind}"] = cat(*(s + f"attn.{n}.{kind}" for n in ("add_q_proj", "add_k_proj", "add_v_proj")))
        t[p + "img_attn.norm.query_norm.weight"] = sd[s + "attn.norm_q.weight"]
        t[p + "img_attn.norm.key_norm.weight"] = sd[s + "attn.norm_k.weight"]
        t[p + "txt_attn.norm.query_norm.weight"] = sd[s + "attn.norm_added_q.weight"]
        t[p + "txt_attn.norm.key_norm.weight"] = sd[s + "attn.norm_added_k.weight"]
        lin(p + "img_attn.proj", s + "attn.to_out.0")
        lin(p + "txt_attn.proj", s + "attn.to_add_out")
        lin(p + "img_mlp.0", s + "ff.net.0.proj")
        lin(p + "img_mlp.2", s + "ff.net.2")
        lin(p + "txt_mlp.0", s + "ff_context.net.0.proj")
        lin(p + "txt_mlp.2", s + "ff_context.net.2")
    for i in range(cfg.num_single_layers):
        s, p = f"single_transformer_blocks.{i}.", f"single_blocks.{i}."
        lin(p + "modulation.lin", s + "norm.linear")
        for kind in ("weight", "bias"):
            t[p + f"linear1.{kind}"] = cat(*(s + f"attn.{n}.{kind}" for n in ("to_q", "to_k", "to_v")), s + f"proj_mlp.{kind}")
        lin(p + "linear2", s + "proj_out")
        t[p + "norm.query_norm.weight"] = sd[s + "attn.norm_q.weight"]
        t[p + "norm.key_norm.weight"] = sd[s + "attn.norm_k.weight"]
    lin("final_layer.linear", "proj_out")
    # BFL chunks (shift, scale); diffusers AdaLayerNormContinuous chunks (scale, shift): swap halves
    for kind in ("weight", "bias"):
        a = sd[f"norm_out.linear.{kind}"]
        t[f"final_layer.adaLN_modulation.1.{kind}"] = torch.cat([a[d:], a[:d]], dim=0)
    missing, unexpected = m.load_state_dict({k: v.double() for k, v in t.items()}, strict=True)
    for mod in m.modules():
        if isinstance(mod, torch.nn.RMSNorm):
            mod.eps = 1e-6  # diffusers RMSNorm eps (torchtitan default is finfo.eps)
    return m.eval()


def titan_forward_with_guidance(m, cfg: fo.FluxConfig, sd: dict, inp: dict):
    """torchtitan's FluxModel has no guidance embedder (it implements FLUX.1-schnell's graph); FLUX.1-dev / Kontext add
    `vec += guidance_in(timestep_embedding(guidance, 256))` with `guidance_in` an MLPEmbedder like `time_in` (BFL
    model.py).  This bolts that term onto torchtitan's own modules — its MLPEmbedder class and its
    timestep_embedding — and otherwise replays FluxModel.forward verbatim, so the oracle's guidance MLP and the place it
    enters the graph are checked by an implementation this repo did not write."""
    from torchtitan.experiments.flux.model.layers import MLPEmbedder, timestep_embedding

    g_in = MLPEmbedder(in_dim=256, hidden_dim=cfg.inner_dim).double()
    g_in.load_state_dict({"in_layer.weight": sd["time_text_embed.guidance_embedder.linear_1.weight"].double(),
                          "in_layer.bias": sd["time_text_embed.guidance_embedder.linear_1.bias"].double(),
                          "out_layer.weight": sd["time_text_embed.guidance_embedder.linear_2.weight"].double(),
                          "out_layer.bias": sd["time_text_embed.guidance_embedder.linear_2.bias"].double()})
    img, txt = m.img_in(inp["hidden_states"]), m.txt_in(inp["encoder_hidden_states"])
    vec = m.time_in(timestep_embedding(inp["timestep"], 256))
    vec = vec + g_in(timestep_embedding(inp["guidance"], 256))
    vec = vec + m.vector_in(inp["pooled_projections"])
    ids = torch.cat((inp["txt_ids"][None].double(), inp["img_ids"][None].double()), dim=1)
    ids = ids.expand(img.shape[0], -1, -1)
    pe = m.pe_embedder(ids)
    for block in m.double_blocks:
        img, txt = block(img=img, txt=txt, vec=vec, pe=pe)
    img = torch.cat((txt, img), 1)
    for block in m.single_blocks:
        img = block(img, vec=vec, pe=pe)
    img = img[:, txt.shape[1]:, ...]
    return m.final_layer(img, vec)


def titan_guidance_golden():
    """flux_toy_titan_guidance.pt: torchtitan + the bolted-on guidance embedder, float64, guidance 3.5 and 1.0."""
    cfg = fo.FluxConfig(**TOY)                      # guidance_embeds=True
    sd = fo.make_synthetic_state_dict(cfg, seed=13, dtype=torch.float64)
    inp = toy_inputs(cfg, B=2, seed=5)
    inp["guidance"] = torch.tensor([3.5, 1.0], dtype=torch.float64)
    m = titan_from_diffusers(cfg, {k: v for k, v in sd.items() if "guidance_embedder" not in k})
    with torch.no_grad():
        out = titan_forward_with_guidance(m, cfg, sd, inp)
    torch.save(dict(cfg=TOY, seed=13, inputs={k: v.float() for k, v in inp.items()}, output=out.float()),
               HERE / "flux_toy_titan_guidance.pt")
    print("titan+guidance output", out.shape, out.abs().mean().item())


VAE_TOY = dict(block_out_channels=(32, 64, 128, 128))


def titan_ae_from_diffusers(vcfg, sd):
    """torchtitan AutoEncoder (BFL layout) loaded with diffusers-named VAE weights."""
    from torchtitan.experiments.flux.model.autoencoder import AutoEncoder, AutoEncoderParams

    boc = vcfg.block_out_channels
    ch = boc[0]
    params = AutoEncoderParams(resolution=64, in_channels=3, ch=ch, out_ch=3, ch_mult=[b // ch for b in boc],
                               num_res_blocks=vcfg.layers_per_block, z_channels=vcfg.latent_channels,
                               scale_factor=vcfg.scaling_factor, shift_factor=vcfg.shift_factor)
    ae = AutoEncoder(params).double()
    t = {}
    n = len(boc)

    def cp(dst, src):
        for kind in ("weight", "bias"):
            t[f"{dst}.{kind}"] = sd[f"{src}.{kind}"]

    def res(dst, src):
        for a in ("norm1", "conv1", "norm2", "conv2"):
            cp(f"{dst}.{a}", f"{src}.{a}")
        if f"{src}.conv_shortcut.weight" in sd:
            cp(f"{dst}.nin_shortcut", f"{src}.conv_shortcut")

    def mid(dst, src):
        res(f"{dst}.block_1", f"{src}.resnets.0")
        res(f"{dst}.block_2", f"{src}.resnets.1")
        cp(f"{dst}.attn_1.norm", f"{src}.attentions.0.group_norm")
        for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
            t[f"{dst}.attn_1.{a}.weight"] = sd[f"{src}.attentions.0.{b}.weight"][:, :, None, None]
            t[f"{dst}.attn_1.{a}.bias"] = sd[f"{src}.attentions.0.{b}.bias"]

    cp("encoder.conv_in", "encoder.conv_in")
    for i in range(n):
        for j in range(vcfg.layers_per_block):
            res(f"encoder.down.{i}.block.{j}", f"encoder.down_blocks.{i}.resnets.{j}")
        if i != n - 1:
            cp(f"encoder.down.{i}.downsample.conv", f"encoder.down_blocks.{i}.downsamplers.0.conv")
    mid("encoder.mid", "encoder.mid_block")
    cp("encoder.norm_out", "encoder.conv_norm_out")
    cp("encoder.conv_out", "encoder.conv_out")
    cp("decoder.conv_in", "decoder.conv_in")
    mid("decoder.mid", "decoder.mid_block")
    for i in range(n):  # diffusers up_blocks[0] is the lowest resolution = BFL up[n-1]
        for j in range(vcfg.layers_per_block + 1):
            res(f"decoder.up.{n - 1 - i}.block.{j}", f"decoder.up_blocks.{i}.resnets.{j}")
        if i != n - 1:
            cp(f"decoder.up.{n - 1 - i}.upsample.conv", f"decoder.up_blocks.{i}.upsamplers.0.conv")
    cp("decoder.norm_out", "decoder.conv_norm_out")
    cp("decoder.conv_out", "decoder.conv_out")
    ae.load_state_dict({k: v.double() for k, v in t.items()}, strict=True)
    return ae.eval()


def vae_golden():
    from oracle import vae_oracle as vo

    vcfg = vo.VaeConfig(**VAE_TOY)
    sd = vo.make_synthetic_state_dict(vcfg, seed=21, dtype=torch.float64)
    g = torch.Generator().manual_seed(22)
    x = torch.rand(1, 3, 48, 64, generator=g, dtype=torch.float64) * 2 - 1
    z = torch.randn(1, 16, 6, 8, generator=g, dtype=torch.float64)
    ae = titan_ae_from_diffusers(vcfg, sd)
    with torch.no_grad():
        moments = ae.encoder(x)
        mean = torch.chunk(moments, 2, dim=1)[0]
        img = ae.decoder(z)
    torch.save(dict(cfg=VAE_TOY, seed=21, x=x.float(), z=z.float(), mean=mean.float(), image=img.float()),
               HERE / "vae_toy_titan.pt")
    print("vae titan", mean.shape, mean.abs().mean().item(), img.shape, img.abs().mean().item())


def main():
    vae_golden()
    rope_index_golden()
    titan_guidance_golden()
    cfg = fo.FluxConfig(**TOY, guidance_embeds=False)
    sd = fo.make_synthetic_state_dict(cfg, seed=11, dtype=torch.float64)
    inp = toy_inputs(cfg)
    m = titan_from_diffusers(cfg, sd)
    with torch.no_grad():
        out = m(img=inp["hidden_states"], img_ids=inp["img_ids"][None].double(), txt=inp["encoder_hidden_states"],
                txt_ids=inp["txt_ids"][None].double(), timesteps=inp["timestep"], y=inp["pooled_projections"])
    torch.save(dict(cfg=TOY, seed=11, inputs={k: v.float() for k, v in inp.items()}, output=out.float()),
               HERE / "flux_toy_titan.pt")
    print("titan output", out.shape, out.abs().mean().item())

    cfg2 = fo.FluxConfig(**TOY)
    sd2 = fo.make_synthetic_state_dict(cfg2, seed=12, dtype=torch.float32)
    inp2 = toy_inputs(cfg2, B=2, seed=2, dtype=torch.float32)
    inp2["timestep"] = (torch.tensor([988.4086, 500.0]).bfloat16() / 1000)
    with torch.no_grad():
        out2 = fo.flux_forward(sd2, cfg2, inp2["hidden_states"], inp2["encoder_hidden_states"], inp2["pooled_projections"],
                               inp2["timestep"], inp2["img_ids"], inp2["txt_ids"], guidance=inp2["guidance"].float())
    torch.save(dict(cfg=TOY, seed=12, inputs=inp2, output=out2), HERE / "flux_toy_oracle.pt")
    print("oracle output", out2.shape, out2.abs().mean().item())



def rope_index_golden():
    """Runs the REFERENCE's own get_rope_index (source text of
    univa/models/qwen2p5vl/modeling_univa_qwen2p5vl.py:139-318, extracted with ast because the module
    itself needs diffusers/transformers-4.50 to import) on a few token layouts -> rope_index_ref.pt."""
    import ast
    import textwrap
    from types import SimpleNamespace
    from typing import Optional, Tuple  # noqa: F401  (names used by the extracted source)

    src = Path("/root/reference/univa/models/qwen2p5vl/modeling_univa_qwen2p5vl.py").read_text()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "get_rope_index")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    ns = {"torch": torch, "Optional": Optional, "Tuple": Tuple}
    exec(code, ns)
    fake = SimpleNamespace(config=SimpleNamespace(vision_config=SimpleNamespace(spatial_merge_size=2, tokens_per_second=2),
                                                   image_token_id=900, video_token_id=901, vision_start_token_id=902))
    img32, img4, img256 = [900] * 32, [900] * 4, [900] * 256
    cases = [
        (torch.tensor([[1, 2, 3, 902] + img32 + [903, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14],
                       [1, 902] + img4 + [903, 4, 902] + img32 + [903, 5, 6, 7, 8, 9]]),
         torch.tensor([[1, 16, 8], [1, 4, 4], [1, 16, 8]])),
        (torch.tensor([[1, 2, 3, 902] + img256 + [903] + list(range(10, 37))]), torch.tensor([[1, 32, 32]])),
        (torch.tensor([[7, 8, 902] + img4 + [903, 9, 902]]), torch.tensor([[1, 4, 4]])),   # trailing vision_start
    ]
    out = []
    for ids, grid in cases:
        pos, delta = ns["get_rope_index"](fake, ids, grid)
        out.append(dict(input_ids=ids, image_grid_thw=grid, position_ids=pos, deltas=delta))
    # padded batches (processor(..., padding=True)): right padding, left padding, text only
    pad = 0
    row_a = [1, 2, 3, 902] + img32 + [903, 5, 6, 7, 8, 9, 10]
    row_b = [1, 902] + img4 + [903, 4, 5]
    n = len(row_a)
    masked = [
        (torch.tensor([row_a, row_b + [pad] * (n - len(row_b))]),
         torch.tensor([[1] * n, [1] * len(row_b) + [0] * (n - len(row_b))]), torch.tensor([[1, 16, 8], [1, 4, 4]])),
        (torch.tensor([row_a, [pad] * (n - len(row_b)) + row_b]),
         torch.tensor([[1] * n, [0] * (n - len(row_b)) + [1] * len(row_b)]), torch.tensor([[1, 16, 8], [1, 4, 4]])),
        (torch.tensor([[5, 6, 7, 8, 9, 10], [5, 6, 7, pad, pad, pad]]), torch.tensor([[1] * 6, [1, 1, 1, 0, 0, 0]]), None),
    ]
    for ids, mask, grid in masked:
        pos, delta = ns["get_rope_index"](fake, ids, grid, None, None, mask)
        out.append(dict(input_ids=ids, image_grid_thw=grid, attention_mask=mask, position_ids=pos, deltas=delta))
    torch.save(out, HERE / "rope_index_ref.pt")
    print("rope_index golden:", [tuple(o["position_ids"].shape) for o in out])
    # a randomized sweep: 60 batches of 1-3 prompts with 0-3 images each (even grids up to 12x16 patches), random text
    # runs between them, padded to the longest prompt on a random side
    import random
    rng = random.Random(1234)
    sweep = []
    for _ in range(60):
        rows, grids = [], []
        for _b in range(rng.randint(1, 3)):
            toks = [rng.randint(1, 800) for _ in range(rng.randint(0, 6))]
            for _i in range(rng.randint(0, 3)):
                gh, gw = 2 * rng.randint(1, 6), 2 * rng.randint(1, 8)
                grids.append([1, gh, gw])
                toks += [902] + [900] * (gh * gw // 4) + [903] + [rng.randint(1, 800) for _ in range(rng.randint(0, 9))]
            if not toks:
                toks = [5]
            rows.append(toks)
        n = max(map(len, rows))
        left = rng.random() < 0.5
        ids = torch.tensor([([0] * (n - len(r)) + r) if left else (r + [0] * (n - len(r))) for r in rows])
        mask = torch.tensor([([0] * (n - len(r)) + [1] * len(r)) if left else ([1] * len(r) + [0] * (n - len(r))) for r in rows])
        grid = torch.tensor(grids) if grids else None
        use_mask = not bool(mask.all()) or rng.random() < 0.5
        pos, delta = ns["get_rope_index"](fake, ids, grid, None, None, mask if use_mask else None)
        sweep.append(dict(input_ids=ids.to(torch.int16), image_grid_thw=grid, attention_mask=mask.to(torch.int8) if use_mask else None,
                          position_ids=pos.to(torch.int16), deltas=delta.to(torch.int16)))
    torch.save(sweep, HERE / "rope_index_sweep_ref.pt")
    print("rope_index sweep:", len(sweep), "cases,", sum(c["image_grid_thw"] is not None for c in sweep), "with images,",
          sum(c["attention_mask"] is not None for c in sweep), "with a mask")


if __name__ == "__main__":
    main()
