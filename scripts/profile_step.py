"""Short workload for ncu: one warm-up + one profiled MMDiT forward at C1024 (B=1, S=8736) and,
optionally, one VAE encode/decode.  All libb2f kernels live in namespace b2f (-k regex:b2f)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="19,38")
ap.add_argument("--forwards", type=int, default=2)
ap.add_argument("--vae", action="store_true")
ap.add_argument("--res", type=int, default=1024)
a = ap.parse_args()
nd, ns = map(int, a.layers.split(","))
dev = torch.device("cuda")
m = B200FluxTransformer2DModel(FluxTransformerConfig(num_layers=nd, num_single_layers=ns)).randomize_(0)
g = torch.Generator(device=dev).manual_seed(0)
n = (a.res // 16) ** 2
hs = torch.randn(1, 2 * n, 64, device=dev, generator=g).bfloat16()
enc = torch.randn(1, 544, 4096, device=dev, generator=g).bfloat16()
pooled = torch.randn(1, 768, device=dev, generator=g).bfloat16()
ids = torch.zeros(2 * n, 3, device=dev, dtype=torch.bfloat16)
txt = torch.zeros(544, 3, device=dev, dtype=torch.bfloat16)
t = torch.full((1,), 0.5, device=dev).bfloat16()
gd = torch.full((1,), 3.5, device=dev)
for _ in range(a.forwards):
    m(hidden_states=hs, encoder_hidden_states=enc, pooled_projections=pooled, timestep=t, img_ids=ids, txt_ids=txt,
      guidance=gd, return_dict=False)
torch.cuda.synchronize()
if a.vae:
    from gpt_image_edit_b200.vae import B200AutoencoderKL
    v = B200AutoencoderKL().randomize_(1)
    img = (torch.rand(1, 3, a.res, a.res, device=dev, generator=g) * 2 - 1).bfloat16()
    z = v.encode(img).latent_dist.mode()
    v.decode(z, return_dict=False)
    torch.cuda.synchronize()
print("done")
