"""The prompt path of `univa.serve.cli` / the GEdit driver against the reference's OWN statements (cli.py:151-197,
gedit/step1_gen_samples.py:97-151), executed by tests/golden/make_cli_chat_golden.py into cli_chat_ref.pt:
the user's instruction must reach the VLM through the chat template — same `input_ids`, `attention_mask`,
`image_grid_thw` and `pixel_values` as the reference's code produces for the same conversation, turn after turn.
"""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "cli_chat_ref.pt"
sys.path.insert(0, str(Path(__file__).parent))


@pytest.fixture(scope="module")
def env(tmp_path_factory):
    from PIL import Image
    from transformers import AutoProcessor

    from toy_processor import build_toy_processor

    gold = torch.load(GOLD, weights_only=False)
    td = tmp_path_factory.mktemp("chat")
    rng = np.random.default_rng(gold["image_seed"])
    for name, shape in gold["image_shapes"].items():
        arr = rng.integers(0, 256, size=shape, dtype=np.uint8)
        assert int(arr.astype(np.int64).sum()) == gold["image_sums"][name], "numpy RNG stream changed: regenerate the fixture"
        Image.fromarray(arr).save(td / name)
    build_toy_processor(td / "proc")
    from gpt_image_edit_b200.checkpoint import load_processor
    processor = load_processor(td / "proc", 448 * 448, 448 * 448)
    assert type(processor).__name__ == type(AutoProcessor.from_pretrained(str(td / "proc"))).__name__
    return SimpleNamespace(gold=gold, td=td, processor=processor)


def _check(rec, ids, mask, pix, grid):
    assert torch.equal(ids.cpu(), rec["input_ids"])
    assert torch.equal(mask.cpu(), rec["attention_mask"])
    if "image_grid_thw" in rec:
        assert torch.equal(grid.cpu(), rec["image_grid_thw"])
        assert tuple(pix.shape) == rec["pixel_values_shape"]
        assert abs(pix.double().sum().item() - rec["pixel_values_sum"]) < 1e-6 * max(1.0, abs(rec["pixel_values_abs_sum"]))
        assert torch.equal(pix[:2, :16].cpu(), rec["pixel_values_head"])
    else:
        assert pix is None and grid is None


def test_chat_session_builds_the_references_inputs_turn_after_turn(env):
    from univa.serve import cli

    args = SimpleNamespace(height=1024, width=1024, ocr_enhancer=False, no_auto_hw=False, synthetic=False)
    sess = cli.ChatSession(args, None, None, None, env.processor, None, None, "cpu")
    for rec in env.gold["cli_turns"]:
        urls = [str(env.td / rec["image"])] if rec["image"] else []
        new_h, new_w = sess.add_user_turn(rec["txt"], urls)
        assert (new_h, new_w) == (rec["new_h"], rec["new_w"])
        assert len(sess.history_image_paths) == rec["n_history"]
        _check(rec, *sess.model_inputs())
        # the instruction text itself is in the prompt the VLM sees
        text = env.processor.batch_decode(sess.model_inputs()[0], skip_special_tokens=True, clean_up_tokenization_spaces=False)[0]
        assert rec["txt"] in text and "You are a helpful assistant" not in text
        sess.conversation.append({"role": "assistant", "content": [{"type": "text", "text": "ok " + rec["txt"][:5]}]})


def test_two_images_in_one_turn(env):
    from univa.serve import cli

    args = SimpleNamespace(height=1024, width=1024, ocr_enhancer=False, no_auto_hw=False, synthetic=False)
    sess = cli.ChatSession(args, None, None, None, env.processor, None, None, "cpu")
    rec = env.gold["cli_two_images"]
    new_h, new_w = sess.add_user_turn("blend them", [str(env.td / "a.png"), str(env.td / "b.png")])
    assert (new_h, new_w) == (rec["new_h"], rec["new_w"])
    _check(rec, *sess.model_inputs())


def test_gedit_prompt_path(env):
    from univa.eval.gedit.step1_gen_samples import generation_size
    from univa.serve import cli

    rec = env.gold["gedit"]
    content = [{"type": "image", "image": str(env.td / rec["image"]), "resized_height": 448, "resized_width": 448},
               {"type": "text", "text": rec["prompt"]}]
    inputs = cli.prepare_inputs(env.processor, [{"role": "user", "content": content}], "cpu")
    _check(rec, inputs["input_ids"], inputs["attention_mask"], inputs["pixel_values"], inputs["image_grid_thw"])
    h, w, _ = env.gold["image_shapes"][rec["image"]]
    assert generation_size(h, w, 1024, 1024) == (rec["gen_h"], rec["gen_w"])


def test_no_processor_is_an_error_unless_synthetic(env, tmp_path):
    from gpt_image_edit_b200._lib import B2FError
    from gpt_image_edit_b200.checkpoint import load_processor
    from univa.serve import cli

    with pytest.raises(FileNotFoundError):
        load_processor(tmp_path)
    args = SimpleNamespace(height=1024, width=1024, synthetic=False)
    with pytest.raises(B2FError):
        cli.ChatSession(args, None, None, None, None, None, None, "cpu")
    cli.ChatSession(SimpleNamespace(height=1024, width=1024, synthetic=True), None, None, None, None, None, None, "cpu")


def test_process_vision_info_matches_the_image_processors_own_resize(env):
    """the restated `qwen_vl_utils.smart_resize` agrees with transformers' (the image processor applies it again)."""
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import smart_resize as hf

    from gpt_image_edit_b200.image_io import process_vision_info, smart_resize

    rng = np.random.default_rng(0)
    for _ in range(2000):
        h, w = int(rng.integers(20, 3000)), int(rng.integers(20, 3000))
        if max(h, w) / min(h, w) > 150:
            continue
        for lo, hi in ((448 * 448, 448 * 448), (4 * 28 * 28, 16384 * 28 * 28)):
            assert smart_resize(h, w, 28, lo, hi) == hf(h, w, 28, lo, hi)
    imgs, vids = process_vision_info([{"role": "user", "content": [
        {"type": "image", "image": str(env.td / "a.png"), "min_pixels": 448 * 448, "max_pixels": 448 * 448}]}])
    assert vids is None and imgs[0].size == (532, 392)            # 420x300 scaled UP to the 448^2 budget (ceil), aspect kept; the image processor then floors to 504x364
    assert process_vision_info([{"role": "user", "content": [{"type": "text", "text": "hi"}]}]) == (None, None)


def test_fetch_image_sources_of_qwen_vl_utils(tmp_path):
    """qwen_vl_utils.fetch_image takes a PIL image, a local path, `file://`, a base64 `data:image` URI (and http urls, which
    need a network): all must give the same resized RGB view."""
    import base64
    import io

    import numpy as np
    from PIL import Image

    from gpt_image_edit_b200.image_io import fetch_image

    im = Image.fromarray(np.random.default_rng(0).integers(0, 256, (60, 90, 3), dtype=np.uint8))
    path = tmp_path / "a.png"
    im.save(path)
    buf = io.BytesIO()
    im.save(buf, format="PNG")
    budget = {"min_pixels": 448 * 448, "max_pixels": 448 * 448}
    views = [fetch_image({"type": "image", "image": src, **budget}) for src in
             (im, str(path), "file://" + str(path), "data:image;base64," + base64.b64encode(buf.getvalue()).decode())]
    assert views[0].size == (560, 392) and views[0].mode == "RGB"                 # 28-multiples, ~448*448 pixels, aspect kept
    assert all(np.array_equal(np.asarray(views[0]), np.asarray(v)) for v in views[1:])
    with pytest.raises(ValueError, match="Unrecognized image input"):
        fetch_image({"type": "image", "image": "data:image/png,not-base64"})
