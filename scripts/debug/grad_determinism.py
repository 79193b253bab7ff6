"""1-GPU diagnostic: are the training gradients bit-reproducible at full width (pair GEMM kernels), one call vs block by block?"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def main():
    from univa.training.configuration_denoise import from_mapping
    from univa.training.synthetic_data import SyntheticEditDataset, collate
    import train_denoiser as td
    from gpt_image_edit_b200.training import Stage2Trainer, FluxTrainGraph, flow_matching_loss

    dev = torch.device("cuda")
    conf = from_mapping(dict(
        training_config=dict(seed=5, learning_rate=1e-4, adam_beta2=0.99, adam_weight_decay=0.0, discrete_timestep=False,
                             mask_weight_type="log", max_grad_norm=1e9),
        model_config=dict(synthetic=True, small=True, with_tune_mlp2=True, joint_ref_feature=True,
                          flux_train_layer_idx=list(range(57))),
        dataset_config=dict(dataset_type="synthetic", batch_size=1, height=256, width=256)))
    model, vae, pipe, empty = td.build_models(conf, dev)
    tr = Stage2Trainer(model, vae, pipe, conf.training_config, conf.model_config, empty)
    data = SyntheticEditDataset(256, 256, seed=1)
    batch = collate([data[0]])
    grads = []
    for mode in ("one", "one", "blocks", "blocks"):
        tr.graph.on_block_done = (lambda b: None) if mode == "blocks" else None
        tr.opt.step = lambda lr=None: torch.zeros(1, device=dev)      # no update: same weights every time
        tr.gen = torch.Generator(device=dev).manual_seed(100)
        out = tr.step(batch)
        torch.cuda.synchronize()
        grads.append((mode, out["loss"].item(), [p.grad.clone() for p in tr.params]))
    base = grads[0]
    for mode, loss, g in grads[1:]:
        bad = []
        for p, a, b in zip(tr.params, base[2], g):
            if not torch.equal(a, b):
                d = (a - b).abs()
                bad.append((p.name, int((d > 0).sum()), float(d.max()), float(a.abs().max())))
        print(mode, "loss", loss, "vs", base[1], "differing tensors:", bad)


if __name__ == "__main__":
    main()
