"""GPU smoke of the reference-facing entry point: one non-interactive `univa.serve.cli` turn with
synthetic weights (few layers at the real widths), 256x256 / 2 steps — BASELINE.json configs[0]'s
plumbing case, run on the GPU because this engine has no CPU path — plus the checkpoint round trip."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cli_single_turn_writes_png(tmp_path):
    from PIL import Image

    from univa.serve import cli

    rng = np.random.default_rng(0)
    src = tmp_path / "src.png"
    Image.fromarray(rng.integers(0, 256, size=(300, 420, 3), dtype=np.uint8)).save(src)
    out = tmp_path / "out.png"
    args = cli.build_parser().parse_args([
        "--synthetic", "--small", "--prompt", "make the sky purple", "--image", str(src), "--output", str(out),
        "--height", "256", "--width", "256", "--num_inference_steps", "2", "--max_area", str(256 * 256)])
    cli.main(args)
    img = Image.open(out)
    # update_size(any_11ratio, anchor 256*256) of a 420x300 image, then the pipeline's /16 floor
    assert img.size[0] % 16 == 0 and img.size[1] % 16 == 0 and img.size[0] > img.size[1]
    assert np.asarray(img).std() > 0


def test_cli_text_reply_branch(tmp_path, capsys):
    """The understanding branch (reference cli.py:256-267): prefill -> task head -> greedy KV-cache decode."""
    from PIL import Image

    from univa.serve import cli

    rng = np.random.default_rng(1)
    src = tmp_path / "src.png"
    Image.fromarray(rng.integers(0, 256, size=(224, 224, 3), dtype=np.uint8)).save(src)
    args = cli.build_parser().parse_args(["--synthetic", "--small", "--prompt", "what is in the picture?", "--image", str(src),
                                          "--force_text_reply", "--max_new_tokens", "6"])
    cli.main(args)
    out = capsys.readouterr().out
    assert "Assistant: <token ids>" in out
    toks = out.split("<token ids>")[1].split()
    assert 1 <= len(toks) <= 6 and all(0 <= int(t) < 152064 for t in toks)


def test_cli_turn_through_a_processor_the_instruction_reaches_the_vlm(tmp_path, monkeypatch):
    """Non-synthetic prompt path on the GPU: chat template + tokenizer + image processor (a toy Qwen2.5-VL processor
    directory) feed the prefill; two different instructions give different VLM embeddings and the turn writes a PNG."""
    import sys
    from pathlib import Path
    from types import SimpleNamespace

    from PIL import Image

    sys.path.insert(0, str(Path(__file__).parent))
    from toy_processor import build_toy_processor

    from gpt_image_edit_b200.checkpoint import load_processor
    from univa.serve import cli

    build_toy_processor(tmp_path / "proc")
    processor = load_processor(tmp_path / "proc")
    tid = processor.tokenizer.convert_tokens_to_ids
    dev = torch.device("cuda")
    model, head, _ = cli.load_main_model_and_processor("", dev, synthetic=True, small=True, image_token_id=tid("<|image_pad|>"),
                                                       vision_start_token_id=tid("<|vision_start|>"))
    pipe, toks, encs = cli.load_pipe(model.denoise_tower.denoiser, "", dev, synthetic=True, small=True)
    monkeypatch.setattr(cli, "ASSISTANT_TOKEN_ID", tid("<|im_start|>"))   # the toy vocabulary has no 77091
    head.w3.zero_()                                                       # logits = bias = (0, 1): the turn generates
    rng = np.random.default_rng(3)
    src = tmp_path / "src.png"
    Image.fromarray(rng.integers(0, 256, size=(300, 420, 3), dtype=np.uint8)).save(src)
    args = cli.build_parser().parse_args(["--height", "256", "--width", "256", "--num_inference_steps", "2",
                                          "--max_area", str(256 * 256)])
    embeds = []
    for text in ("make the sky purple", "remove the tree on the left"):
        sess = cli.ChatSession(args, model, head, pipe, processor, toks, encs, dev)
        sess.add_user_turn(text, [str(src)])
        ids, mask, pix, grid = sess.model_inputs()
        assert grid.tolist() == [[1, 26, 36]] and int((ids == tid("<|image_pad|>")).sum()) == 26 * 36 // 4
        embeds.append(model(ids, pixel_values=pix, attention_mask=mask, image_grid_thw=grid, output_type="denoise_embeds"))
    assert embeds[0].shape[1] != embeds[1].shape[1] or not torch.equal(embeds[0], embeds[1])
    sess = cli.ChatSession(args, model, head, pipe, processor, toks, encs, dev)
    kind, out = sess.turn("make the sky purple", [str(src)], output_path=str(tmp_path / "out.png"))
    assert kind == "image" and Image.open(out).size[0] % 16 == 0
    assert sess.conversation[-1] == {"role": "assistant", "content": [{"type": "image", "image": out}]}
    assert sess.history_image_paths == [str(src), out]


def test_checkpoint_roundtrip_through_safetensors(tmp_path):
    from gpt_image_edit_b200 import checkpoint as ck
    from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig

    cfg = dict(num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=2, joint_attention_dim=256,
               pooled_projection_dim=64)
    m = B200FluxTransformer2DModel(FluxTransformerConfig(**cfg)).randomize_(3)
    ck.save_state_dict(m.state_dict(), tmp_path / "transformer")
    (tmp_path / "transformer" / "config.json").write_text(__import__("json").dumps(cfg))
    m2 = ck.load_flux_transformer(tmp_path / "transformer")
    a, b = m.state_dict(), m2.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert "transformer_blocks.0.attn.to_q.weight" in a and "single_transformer_blocks.0.proj_mlp.bias" in a

    v = B200AutoencoderKL(VaeConfig(block_out_channels=(64, 128, 256, 256))).randomize_(4)
    ck.save_state_dict({k: t.contiguous() for k, t in v.state_dict().items()}, tmp_path / "vae")
    (tmp_path / "vae" / "config.json").write_text('{"block_out_channels": [64, 128, 256, 256]}')
    (tmp_path / "scheduler").mkdir()
    (tmp_path / "scheduler" / "scheduler_config.json").write_text('{"base_shift": 0.5, "max_shift": 1.15}')
    v2, sched = ck.load_pipeline_components(tmp_path)
    sa, sb = v.state_dict(), v2.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa)
    assert sa["encoder.conv_in.weight"].shape == (64, 3, 3, 3)          # diffusers OIHW layout at the boundary
    assert sched.config.max_shift == 1.15


def test_text_encoder_checkpoints_and_string_prompts(tmp_path):
    """FLUX.1 directory layout round trip for text_encoder / text_encoder_2, then a string-prompt pipeline call
    (reference flux_pipeline.py:925-944) that must equal the prompt_embeds-driven call with the same encoders."""
    import json

    from gpt_image_edit_b200 import checkpoint as ck
    from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig
    from gpt_image_edit_b200.pipeline import FluxKontextPipeline
    from gpt_image_edit_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from gpt_image_edit_b200.text_encoders import (B200CLIPTextModel, B200T5Encoder, CLIPTextConfig, SyntheticTokenizer,
                                                   T5EncoderConfig, encode_prompt)
    from gpt_image_edit_b200.vae import B200AutoencoderKL, VaeConfig

    ccfg = dict(vocab_size=200, hidden_size=256, intermediate_size=512, num_hidden_layers=1, num_attention_heads=4)
    tcfg = dict(vocab_size=128, d_model=256, d_kv=64, num_heads=4, d_ff=512, num_layers=1)
    clip = B200CLIPTextModel(CLIPTextConfig(**ccfg)).randomize_(5)
    t5 = B200T5Encoder(T5EncoderConfig(**tcfg)).randomize_(6)
    ck.save_state_dict(clip.state_dict(), tmp_path / "text_encoder", "model.safetensors")
    ck.save_state_dict(t5.state_dict(), tmp_path / "text_encoder_2", "model.safetensors")
    (tmp_path / "text_encoder" / "config.json").write_text(json.dumps(ccfg))
    (tmp_path / "text_encoder_2" / "config.json").write_text(json.dumps(tcfg))
    clip2, tok, t52, tok2 = ck.load_text_encoders(tmp_path)
    assert tok is None and tok2 is None                       # no tokenizer directories were written
    toks = [SyntheticTokenizer.clip(200), SyntheticTokenizer.t5(128)]
    e1, p1 = encode_prompt([clip, t5], toks, "turn the car blue", 32, "cuda")
    e2, p2 = encode_prompt([clip2, t52], toks, "turn the car blue", 32, "cuda")
    assert torch.equal(e1, e2) and torch.equal(p1, p2) and e1.shape == (1, 32, 256) and p1.shape == (1, 256)

    tr = B200FluxTransformer2DModel(FluxTransformerConfig(num_layers=1, num_single_layers=1, attention_head_dim=128,
                                                          num_attention_heads=2, joint_attention_dim=256,
                                                          pooled_projection_dim=256)).randomize_(3)
    vae = B200AutoencoderKL(VaeConfig(block_out_channels=(64, 128, 256, 256))).randomize_(4)
    pipe = FluxKontextPipeline(transformer=tr, vae=vae, scheduler=FlowMatchEulerDiscreteScheduler(), text_encoder=clip2,
                               tokenizer=toks[0], text_encoder_2=t52, tokenizer_2=toks[1])
    noise = torch.randn(1, 64, 64, generator=torch.Generator().manual_seed(42)).bfloat16().cuda()
    kw = dict(height=128, width=128, num_inference_steps=2, max_area=128 * 128, output_type="latent")
    a = pipe(prompt="turn the car blue", max_sequence_length=32, latents=noise.clone(), **kw).images
    b = pipe(prompt_embeds=e1, pooled_prompt_embeds=p1, latents=noise.clone(), **kw).images
    assert torch.equal(a, b)
    with pytest.raises(ValueError, match="Cannot forward both"):
        pipe(prompt="x", prompt_embeds=e1, pooled_prompt_embeds=p1, **kw)
    bare = FluxKontextPipeline(transformer=tr, vae=vae, scheduler=FlowMatchEulerDiscreteScheduler())
    with pytest.raises(ValueError, match="string prompts need"):
        bare(prompt="x", **kw)


def test_gedit_sampling_driver_writes_strided_outputs(tmp_path, monkeypatch):
    """The multi-GPU caller of the path (reference univa/eval/gedit/step1_gen_samples.py): as rank 1 of 2 (no process
    group: the ranks never communicate while sampling) it must produce exactly items 1, 3 and skip existing files."""
    import json

    from PIL import Image

    from univa.eval.configuration_eval import EvalConfig
    from univa.eval.gedit import step1_gen_samples as drv

    rng = np.random.default_rng(3)
    (tmp_path / "imgs" / "en").mkdir(parents=True)
    spec = {}
    for i, (h, w) in enumerate([(200, 300), (256, 256), (300, 200), (240, 320)]):
        Image.fromarray(rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)).save(tmp_path / "imgs" / "en" / f"{i}.png")
        spec[f"k{i}"] = {"prompt": f"make it {i}", "id": f"en/{i}.png"}
    (tmp_path / "gedit.json").write_text(json.dumps(spec))
    cfg = EvalConfig(output_dir=str(tmp_path / "out"), gedit_prompt_path=str(tmp_path / "gedit.json"),
                     gedit_image_dir=str(tmp_path / "imgs"), height=256, width=256, num_inference_steps=2, synthetic=True,
                     small=True, joint_with_t5=True)
    monkeypatch.setattr(drv.D, "env_world", lambda: (2, 1, 0))
    monkeypatch.setattr(drv.D, "init_from_env", lambda **kw: (2, 1, 0))
    assert drv.main(cfg) == 2
    outs = sorted(p.name for p in (tmp_path / "out" / "en").iterdir())
    assert outs == ["1.png", "3.png"]
    im = Image.open(tmp_path / "out" / "en" / "3.png")
    assert im.size[0] % 16 == 0 and im.size[1] % 16 == 0 and im.size[0] > im.size[1]      # 240x320 source -> landscape
    assert drv.main(cfg) == 0                                                             # everything already exists


def test_denoise_tower_glue_and_mlp2_match_the_references_own_module():
    """tests/golden/tower_ref.pt was produced by executing the reference's modeling_univa_denoise_tower.py with a
    recording denoiser (tests/golden/make_tower_ref_golden.py): same forwarded keywords, same [vlm ; prefix] order, same
    zero txt_ids, same dropped arguments; MLP2 = Linear-SiLU-Linear with the reference's weights."""
    from pathlib import Path

    from univa.models.configuration_univa_denoise_tower import UnivaDenoiseTowerConfig
    from univa.models.modeling_univa_denoise_tower import DenoiseProjector, UnivaDenoiseTower

    fx = torch.load(Path(__file__).parent / "golden" / "tower_ref.pt", weights_only=False)
    cfg = UnivaDenoiseTowerConfig(input_hidden_size=64, output_hidden_size=32,
                                  denoiser_config=dict(num_layers=1, num_single_layers=1, attention_head_dim=128,
                                                       num_attention_heads=2, joint_attention_dim=32, pooled_projection_dim=8))
    tower = UnivaDenoiseTower(cfg)
    calls = []

    class Recorder(torch.nn.Module):
        def forward(self, **kw):
            calls.append(kw)
            return (kw["hidden_states"] * 2,)

    tower.denoiser = Recorder()
    i = {k: v.cuda() for k, v in fx["inputs"].items()}
    runs = {
        "vlm_plus_prefix": lambda: tower(i["hs"], i["t"], i["vlm"], i["pooled"], prefix_prompt_embeds=i["t5"], img_ids=i["img_ids"],
                                         guidance=i["guidance"], joint_attention_kwargs={"attention_mask": torch.ones(2, 18)},
                                         enc_attention_mask=torch.ones(2, 5)),
        "vlm_only": lambda: tower(i["hs"], i["t"], i["vlm"], i["pooled"], img_ids=i["img_ids"], guidance=i["guidance"]),
        "prefix_only": lambda: tower(i["hs"], i["t"], None, i["pooled"], prefix_prompt_embeds=i["t5"], img_ids=i["img_ids"],
                                     guidance=i["guidance"]),
    }
    for name, run in runs.items():
        out = run()
        want = fx["cases"][name]
        got = {k: v for k, v in calls[-1].items() if k != "return_dict"}      # ours asks for the tuple form explicitly
        assert sorted(got) == sorted(want["call"]), name
        for k, v in want["call"].items():
            assert torch.equal(got[k].cpu(), v) if torch.is_tensor(v) else got[k] == v, (name, k)
        assert torch.equal(out.cpu(), want["out"])
    pj = fx["projector"]
    assert pj["structure"] == ["Linear", "SiLU", "Linear"]
    assert pj["keys"] == sorted("denoise_projector." + k for k in DenoiseProjector(64, 32).state_dict())
    mlp2 = DenoiseProjector(64, 32)
    mlp2.load_state_dict({k: v.cuda().bfloat16() for k, v in pj["state_dict"].items()})
    y = mlp2(pj["x"].cuda().bfloat16()).float().cpu()
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    ek, et = rel(y, pj["y_fp32"]), rel(pj["y_bf16"].float(), pj["y_fp32"])
    print(f"MLP2 vs the reference module: kernel {ek:.3e}, reference-in-bf16 {et:.3e}")
    assert y.shape == pj["y_fp32"].shape and ek <= 2.0 * et + 2e-3
