import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from gpt_image_edit_b200 import _lib
from gpt_image_edit_b200._lib import check, ptr, stream_ptr

g = torch.Generator(device="cuda").manual_seed(2)
u8 = torch.randint(0, 256, (2, 96, 128, 3), device="cuda", generator=g, dtype=torch.uint8)
ref = ((u8.permute(0, 3, 1, 2).float() / 255.0) - 0.5) / 0.5
ref = ref.contiguous()
a = torch.zeros(2, 96, 128, 64, device="cuda", dtype=torch.bfloat16)
b = torch.zeros_like(a)
check(_lib.lib.b2f_nchw_to_nhwc_pad(ptr(u8), 2, ptr(a), 2, 3, 96, 128, 64, stream_ptr()), "u8")
check(_lib.lib.b2f_nchw_to_nhwc_pad(ptr(ref), 1, ptr(b), 2, 3, 96, 128, 64, stream_ptr()), "f32")
torch.cuda.synchronize()
print("pad kernels equal:", torch.equal(a, b), (a.float() - b.float()).abs().max().item())
print(a[0, 0, :2, :4], b[0, 0, :2, :4], ref[0, :, 0, :2])
# blend scalar semantics
old = torch.randn(64, 256, device="cuda", generator=g).bfloat16()
emb = torch.randn(64, 256, device="cuda", generator=g).bfloat16()
f = 0.3
want = old * (1 - f) + emb * f
def bf(x): return x.to(torch.bfloat16).float()
c1 = (bf(old.float() * bf(torch.tensor(1 - f))) + bf(emb.float() * bf(torch.tensor(f)))).bfloat16()
c2 = (bf(old.float() * float(1 - f)) + bf(emb.float() * float(f))).bfloat16()
c3 = (bf(old.float() * torch.tensor(1 - f, dtype=torch.float32).item()) + bf(emb.float() * torch.tensor(f, dtype=torch.float32).item())).bfloat16()
print("scalar rounded to bf16:", torch.equal(want, c1), " fp64 scalar:", torch.equal(want, c2), " fp32 scalar:", torch.equal(want, c3))
