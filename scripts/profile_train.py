"""Short workload for ncu: one training forward + backward (MLP2 -> FLUX, per-block recompute) at the 512x512 shapes of
BASELINE.json configs[3] (S = 288 + 1024 + 1024, d = 3072) with --layers double,single blocks (default 1,1: every kernel
of the step occurs, ~120 launches).  All libb2f kernels live in namespace b2f (-k regex:b2f)."""
import argparse
import sys
from pathlib import Path
from types import SimpleNamespace

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gpt_image_edit_b200 import training as tr  # noqa: E402
from gpt_image_edit_b200.flux_transformer import B200FluxTransformer2DModel, FluxTransformerConfig  # noqa: E402
from univa.models.modeling_univa_denoise_tower import DenoiseProjector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="1,1")
ap.add_argument("--res", type=int, default=512)
ap.add_argument("--iters", type=int, default=1)
a = ap.parse_args()
nd, ns = map(int, a.layers.split(","))
dev = torch.device("cuda")
den = B200FluxTransformer2DModel(FluxTransformerConfig(num_layers=nd, num_single_layers=ns)).randomize_(0)
proj = DenoiseProjector(3584, 4096)
g = torch.Generator(device=dev).manual_seed(0)
for t in proj.state_dict().values():
    t.copy_((torch.randn(t.shape, device=dev, generator=g) * 0.02).bfloat16())
model = SimpleNamespace(denoise_tower=SimpleNamespace(denoiser=den, denoise_projector=proj))
params = tr.trainable_params(model)
opt = tr.ShardedAdamW(params, lr=1e-6)
graph = tr.FluxTrainGraph(model, params)
n = (a.res // 16) ** 2
x = torch.randn(1, 288, 3584, device=dev, generator=g).bfloat16()
hs = torch.randn(1, 2 * n, 64, device=dev, generator=g).bfloat16()
pooled = torch.randn(1, 768, device=dev, generator=g).bfloat16()
img_ids = torch.zeros(2 * n, 3, device=dev, dtype=torch.bfloat16)
t = torch.full((1,), 0.5, device=dev).bfloat16()
gd = torch.full((1,), 1.0, device=dev)
target = torch.randn(1, n, 64, device=dev, generator=g)
for _ in range(a.iters):
    pred = graph.forward(x, hs, t, gd, pooled, img_ids, n)
    loss, dpred = tr.flow_matching_loss(pred, target)
    graph.backward(dpred)
    opt.step()
torch.cuda.synchronize()
print("done", float(loss))
