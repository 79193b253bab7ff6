"""Debug: is the toy pipeline bit-reproducible run to run, and where does it start to differ?"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
sys.path.insert(0, str(Path(__file__).resolve().parents[2] / "tests"))
from test_pipeline_gpu import _build  # noqa: E402

pipe, _ = _build()
g = torch.Generator().manual_seed(13)
H = W = 128
image = (torch.randint(0, 256, (1, 3, H, W), generator=g).float() / 127.5 - 1.0).cuda()
pe = torch.randn(1, 24, 256, generator=g).bfloat16().cuda()
pooled = torch.randn(1, 64, generator=g).bfloat16().cuda()
noise = torch.randn(1, 64, 64, generator=g).bfloat16().cuda()
kw = dict(image=image, prompt_embeds=pe, pooled_prompt_embeds=pooled, height=H, width=W, guidance_scale=3.5, max_area=H * W,
          _auto_resize=False, output_type="latent")
for steps in (1, 3):
    outs = [pipe(latents=noise.clone(), num_inference_steps=steps, **kw).images for _ in range(4)]
    print("steps", steps, "equal to run 0:", [bool(torch.equal(outs[0], o)) for o in outs[1:]],
          [float((outs[0].float() - o.float()).abs().max()) for o in outs[1:]])
z = [pipe.vae.encode(image.bfloat16()).latent_dist.mode() for _ in range(3)]
print("vae encode reproducible:", [bool(torch.equal(z[0], t)) for t in z[1:]])
tr = pipe.transformer
ids = torch.zeros(128, 3, device="cuda", dtype=torch.bfloat16)
txt = torch.zeros(24, 3, device="cuda", dtype=torch.bfloat16)
hs = torch.randn(1, 128, 64, device="cuda").bfloat16()
t = torch.full((1,), 0.5, device="cuda").bfloat16()
gd = torch.full((1,), 3.5, device="cuda")
o = [tr(hidden_states=hs, encoder_hidden_states=pe, pooled_projections=pooled, timestep=t, img_ids=ids, txt_ids=txt, guidance=gd,
        return_dict=False)[0] for _ in range(3)]
print("transformer forward reproducible:", [bool(torch.equal(o[0], x)) for x in o[1:]])
