"""CPU tests of the Qwen2.5-VL host bookkeeping (M-RoPE position ids, ViT window order) against the
transformers implementation present in this image (the reference subclasses transformers' classes;
its own get_rope_index — modeling_univa_qwen2p5vl.py:139-318 — is the same algorithm)."""
import pytest
import torch

hf = pytest.importorskip("transformers.models.qwen2_5_vl.modeling_qwen2_5_vl")


def _hf_model():
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLConfig

    cfg = Qwen2_5_VLConfig(
        text_config=dict(hidden_size=256, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1,
                         intermediate_size=512, vocab_size=1000, rope_parameters=dict(rope_type="default", rope_theta=1e6,
                                                                                       mrope_section=[16, 24, 24])),
        vision_config=dict(depth=1, hidden_size=256, num_heads=4, intermediate_size=344, out_hidden_size=256,
                           fullatt_block_indexes=[0]),
        image_token_id=900, video_token_id=901, vision_start_token_id=902, vision_end_token_id=903)
    return hf.Qwen2_5_VLModel(cfg), cfg


def test_get_rope_index_matches_reference_golden():
    """Golden produced by executing the reference's OWN get_rope_index source
    (tests/golden/make_golden.py::rope_index_golden; transformers 5.x changed this function, so it is
    not a valid checker for the 4.50-era reference)."""
    from pathlib import Path

    from gpt_image_edit_b200.qwen2p5vl import get_rope_index

    cases = torch.load(Path(__file__).parent / "golden" / "rope_index_ref.pt")
    assert len(cases) == 6 and sum("attention_mask" in c for c in cases) == 3      # 3 of them padded batches
    for c in cases:
        pos, delta = get_rope_index(c["input_ids"], c["image_grid_thw"], c.get("attention_mask"), image_token_id=900,
                                    vision_start_token_id=902)
        assert torch.equal(pos, c["position_ids"])
        assert torch.equal(delta.flatten(), c["deltas"].flatten())
    # worked example of SURVEY.md Appendix B: 4 prefix tokens get 0..3, a 16x16 llm grid gets t=4, h=4+row, w=4+col
    ids2 = torch.tensor([[1, 2, 3, 902] + [900] * 256 + [903, 7]])
    p2, _ = get_rope_index(ids2, torch.tensor([[1, 32, 32]]), None, image_token_id=900, vision_start_token_id=902)
    assert p2[:, 0, :4].tolist() == [[0, 1, 2, 3]] * 3
    assert p2[:, 0, 4].tolist() == [4, 4, 4] and p2[:, 0, 4 + 17].tolist() == [4, 5, 5]
    assert p2[:, 0, 4 + 256].tolist() == [20, 20, 20]


def test_get_rope_index_matches_the_references_function_on_a_random_sweep():
    """60 random batches (1-3 prompts, 0-3 images each, random text runs, left or right padding, with and without an
    attention mask) through the reference's own get_rope_index (tests/golden/make_golden.py::rope_index_golden ->
    rope_index_sweep_ref.pt): positions and deltas must be identical."""
    from pathlib import Path

    from gpt_image_edit_b200.qwen2p5vl import get_rope_index

    cases = torch.load(Path(__file__).parent / "golden" / "rope_index_sweep_ref.pt")
    assert len(cases) == 60 and sum(c["attention_mask"] is not None for c in cases) >= 20
    for i, c in enumerate(cases):
        mask = None if c["attention_mask"] is None else c["attention_mask"].long()
        pos, delta = get_rope_index(c["input_ids"].long(), c["image_grid_thw"], mask, image_token_id=900, vision_start_token_id=902)
        assert torch.equal(pos, c["position_ids"].long()), i
        assert torch.equal(delta.flatten(), c["deltas"].long().flatten()), i


def test_text_only_positions():
    from gpt_image_edit_b200.qwen2p5vl import get_rope_index

    ids = torch.arange(12).view(2, 6)
    pos, delta = get_rope_index(ids, None, None)
    assert pos.shape == (3, 2, 6) and torch.equal(pos[0, 0], torch.arange(6)) and int(delta.abs().sum()) == 0


def test_window_index_and_rot_pos_match_transformers():
    from gpt_image_edit_b200.qwen2p5vl import vision_rot_pos_ids, vision_window_index

    model, _ = _hf_model()
    for grid in ([[1, 16, 8]], [[1, 32, 32]], [[1, 20, 12], [1, 4, 4]]):
        g = torch.tensor(grid)
        widx_ref, cu_ref = model.visual.get_window_index(g)
        widx, cu = vision_window_index(grid)
        assert torch.equal(widx, widx_ref)
        assert cu == torch.unique_consecutive(torch.tensor(cu_ref)).tolist()
        # rot_pos_emb returns the frequencies gathered at these ids; compare through the table
        pos = vision_rot_pos_ids(grid)
        dim = 256 // 4 // 2
        inv = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float) / dim))
        freqs = torch.outer(torch.arange(int(g[:, 1:].max()), dtype=torch.float), inv)
        assert torch.allclose(freqs[pos].flatten(1), model.visual.rot_pos_emb(g))


def test_qwen_pixel_values_match_transformers_image_processor():
    """Patch layout / normalisation of the image fed to the ViT vs transformers' Qwen2VLImageProcessor."""
    import numpy as np
    ip_mod = pytest.importorskip("transformers.models.qwen2_vl.image_processing_qwen2_vl")
    from gpt_image_edit_b200.image_io import qwen_pixel_values

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(448, 448, 3), dtype=np.uint8)
    proc = ip_mod.Qwen2VLImageProcessor(min_pixels=448 * 448, max_pixels=448 * 448)
    ref = proc(images=[img], return_tensors="pt")
    pv, grid = qwen_pixel_values(img)
    assert grid.tolist() == ref["image_grid_thw"].tolist() == [[1, 32, 32]]
    assert pv.shape == ref["pixel_values"].shape == (1024, 1176)
    assert torch.allclose(pv, ref["pixel_values"].float(), atol=1e-5)


def test_hf_state_dict_round_trips_through_the_kernel_layout():
    """B200Qwen2p5VL.load_state_dict (HF names -> fused / padded kernel layout) and .state_dict (back) on the weights of a toy
    transformers Qwen2_5_VLModel: every tensor comes back bit-identical under its 4.50 name, fused tensors sit where the
    kernels read them (q | k | v rows, gate | up rows, vision heads in 128-wide slots with zero padding)."""
    from types import SimpleNamespace

    from gpt_image_edit_b200.qwen2p5vl import B200Qwen2p5VL, QwenTextConfig, QwenVisionConfig, _pad8

    model, _ = _hf_model()
    tc = QwenTextConfig(hidden_size=256, num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=1, intermediate_size=512,
                        vocab_size=1000, image_token_id=900, video_token_id=901, vision_start_token_id=902)
    vc = QwenVisionConfig(depth=1, hidden_size=256, num_heads=4, intermediate_size=344, out_hidden_size=256, fullatt_block_indexes=(0,))
    sd = {k.replace("language_model.", "model."): v.detach() for k, v in model.state_dict().items()}
    sd["lm_head.weight"] = torch.randn(1000, 256)
    me = SimpleNamespace(tc=tc, vc=vc, HP=B200Qwen2p5VL.HP, vhd=256 // 4, vi=_pad8(344), thd=128, _dev=torch.device("cpu"),
                         W=B200Qwen2p5VL.alloc_weights(tc, vc, torch.device("cpu")))
    res = B200Qwen2p5VL.load_state_dict(me, sd)
    assert res.missing_keys == []
    back = B200Qwen2p5VL.state_dict(me)
    assert set(back) == set(sd), sorted(set(back) ^ set(sd))[:6]
    for k, v in sd.items():
        assert back[k].shape == v.shape, k
        assert torch.equal(back[k], v.to(torch.bfloat16)), k
    W = me.W
    q = sd["model.layers.0.self_attn.q_proj.weight"].bfloat16()
    assert torch.equal(W["model.layers.0.qkv.w"][:256], q) and W["model.layers.0.qkv.w"].shape == (256 + 2 * 128, 256)
    assert torch.equal(W["model.layers.0.gu.w"][512:], sd["model.layers.0.mlp.up_proj.weight"].bfloat16())
    vq = W["visual.blocks.0.qkv.w"].view(3, 4, 128, 256)
    assert float(vq[:, :, 64:].abs().max()) == 0 and float(vq[:, :, :64].abs().max()) > 0      # head_dim 64 in a 128-wide slot
