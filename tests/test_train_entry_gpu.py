"""GPU run of the reference-facing training entry point `train_denoiser.py` (reference train_denoiser.py:1621-1633,
loop :829-1181) with the synthetic triples of BASELINE.json configs[3] on a reduced stack (a few layers at the real
widths): steps run, the loss is finite, checkpoints have the reference's layout (denoise_projector.bin next to the
state, train_denoiser.py:1229-1236) and a run resumed from a checkpoint continues from its step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _conf(tmp_path, **tc):
    from univa.training.configuration_denoise import from_mapping

    base = dict(seed=5, output_dir=str(tmp_path / "out"), max_train_steps=3, checkpointing_steps=2, learning_rate=1e-4,
                adam_beta2=0.99, adam_weight_decay=0.0, gradient_checkpointing=True, drop_t5_rate=1.0, discrete_timestep=False,
                mask_weight_type="log", report_to="none")
    base.update(tc)
    return from_mapping(dict(training_config=base,
                             model_config=dict(synthetic=True, small=True, with_tune_mlp2=True, joint_ref_feature=True,
                                               flux_train_layer_idx=list(range(57))),
                             dataset_config=dict(dataset_type="synthetic", batch_size=1, height=256, width=256)))


def test_train_denoiser_entry_steps_checkpoints_and_resumes(tmp_path, capsys):
    import train_denoiser as td

    conf = _conf(tmp_path)
    trainer = td.main(conf)
    out = capsys.readouterr().out
    assert trainer.global_step == 3 and "step 3  loss" in out and "Saved state to" in out
    losses = [float(l.split("loss ")[1].split()[0]) for l in out.splitlines() if l.startswith("step ")]
    assert len(losses) == 3 and all(l == l and 0 < l < 100 for l in losses)
    ck = tmp_path / "out" / "checkpoint-2"
    assert (ck / "denoise_projector.bin").exists() and (ck / "optimizer_rank0.pt").exists() and (ck / "denoiser_trainable").is_dir()
    proj = torch.load(ck / "denoise_projector.bin")
    assert set(proj) == {f"denoise_tower.denoise_projector.{k}" for k in ("0.weight", "0.bias", "2.weight", "2.bias")}
    from gpt_image_edit_b200.checkpoint import load_state_dict_from_dir
    names = set(load_state_dict_from_dir(ck / "denoiser_trainable"))
    assert "transformer_blocks.0.attn.to_q.weight" in names and "single_transformer_blocks.0.norm.linear.bias" in names
    assert not any("ff.net" in n or "add_q_proj" in n for n in names)            # frozen tensors are not in the checkpoint
    assert (ck / "random_states_0.pkl").exists()
    # resume ("latest" = checkpoint-2): optimizer partition, the bf16 weights rounded from its masters, the step counter
    # and the random streams come back; one more step reaches max_train_steps
    conf2 = _conf(tmp_path, resume_from_checkpoint="latest", max_train_steps=3)
    t2 = td.main(conf2)
    out2 = capsys.readouterr().out
    assert "Resuming from checkpoint checkpoint-2" in out2 and out2.count("\nstep ") + out2.startswith("step ") == 1
    assert t2.global_step == 3 and t2.opt.step_count == 3
    loss3 = float(out2.split("step 3  loss ")[1].split()[0])
    assert loss3 == loss3 and 0 < loss3 < 100


def test_gradient_accumulation_and_batch_two(tmp_path, capsys):
    import train_denoiser as td
    from univa.training.configuration_denoise import from_mapping

    conf = _conf(tmp_path, gradient_accumulation_steps=2, max_train_steps=2, checkpointing_steps=1000)
    conf.dataset_config.batch_size = 2
    trainer = td.main(conf)
    assert trainer.global_step == 2 and trainer._micro == 4
    out = capsys.readouterr().out
    assert out.count("\nstep ") + out.startswith("step ") >= 2


def test_mixed_size_batch_masks_the_padding(tmp_path, monkeypatch):
    """A batch whose targets have different sizes (train_denoiser.py:907-916, 1151-1165): latents are zero-padded to the
    largest, the padding carries zero loss weight and the loss is normalised by the unpadded area.  The weights themselves
    are pinned to the reference's statements on CPU (tests/test_host_cpu.py); here the step runs on the engine and the
    kernel's loss is recomputed in torch from what it was given."""
    import train_denoiser as td
    from gpt_image_edit_b200 import training as T

    conf = _conf(tmp_path, max_train_steps=1, checkpointing_steps=1000)
    conf.dataset_config.batch_size = 2
    conf.dataset_config.synthetic_target_sizes = [[256, 256], [192, 320]]
    seen = {}
    real = T.flow_matching_loss

    def spy(pred, target, weight=None, grad_scale=1.0):
        loss, d = real(pred, target, weight=weight, grad_scale=grad_scale)
        seen.update(pred=pred.float().clone(), target=target.clone(), weight=weight.clone(), loss=loss.clone(), d=d.float().clone())
        return loss, d

    monkeypatch.setattr(T, "flow_matching_loss", spy)
    trainer = td.main(conf)
    assert trainer.global_step == 1
    B, S, D = seen["pred"].shape                       # packed [2, (32/2)*(40/2), 64]
    assert (B, S, D) == (2, 16 * 20, 64)
    w = seen["weight"]
    # sample 0 is 32x32 latent inside a 32x40 pad, sample 1 is 24x40: unpack the weights and look at the padding
    wl = w.view(B, 16, 20, 16, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, 16, 32, 40)
    assert float(wl[0, :, :, 32:].abs().max()) == 0 and float(wl[1, :, 24:, :].abs().max()) == 0
    assert float(wl[0, :, :, :32].min()) > 0 and float(wl[1, :, :24, :].min()) > 0
    ref = (w.double() * (seen["pred"].double() - seen["target"].double()) ** 2).mean()
    assert abs(float(seen["loss"]) - float(ref)) <= 1e-4 * abs(float(ref))
    assert float(seen["d"][0].view(16, 20, 16, 2, 2)[:, 16:].abs().max()) == 0      # no gradient into sample 0's padding
    assert torch.isfinite(seen["d"]).all()
