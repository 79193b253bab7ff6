"""Developer micro-benchmark: the attention kernel variants (B2F_ATTN_VARIANT, resolved once per process, so one
child process per variant) on the FLUX joint-attention shapes, burst and power-capped steady state, each checked against
torch SDPA on the same inputs, with cuDNN's SDPA kernel timed beside them.  Run under gpurun:
    python scripts/bench_attn_variants.py --variants 51,53,55 --out gpurun_out/attn_variants.json"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SHAPES = [(1, 24, 8736), (1, 24, 5152), (1, 24, 2592), (4, 24, 8736)]


def child(variant):
    import torch
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "scripts"))
    from bench_kernels import sustained, timeit
    from gpt_image_edit_b200 import ops
    from torch.nn.attention import SDPBackend, sdpa_kernel
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for (B, H, S) in SHAPES:
        torch.manual_seed(S)
        qkv = torch.randn(B, S, 3 * H * 128, device="cuda").bfloat16()
        q = qkv[:, :, : H * 128].unflatten(-1, (H, 128))
        k = qkv[:, :, H * 128: 2 * H * 128].unflatten(-1, (H, 128))
        v = qkv[:, :, 2 * H * 128:].unflatten(-1, (H, 128))
        out = torch.empty(B, S, H * 128, device="cuda", dtype=torch.bfloat16)
        fl = 4.0 * B * H * S * S * 128
        r = dict(variant=variant, B=B, H=H, S=S)
        if variant == "cudnn":
            qt, kt, vt = (x.permute(0, 2, 1, 3).contiguous() for x in (q, k, v))
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                f_ = lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt)
                tb = timeit(f_, iters=10, warmup=3, flush=flush)
                ts = sustained(f_)
        else:
            f_ = lambda: ops.attention(q, k, v, out=out)
            tb = timeit(f_, iters=10, warmup=3, flush=flush)
            ts = sustained(f_)
            if B == 1:
                ref = torch.nn.functional.scaled_dot_product_attention(
                    q.permute(0, 2, 1, 3).float(), k.permute(0, 2, 1, 3).float(), v.permute(0, 2, 1, 3).float())
                ref = ref.permute(0, 2, 1, 3).reshape(B, S, H * 128)
                r["max_abs_err_vs_fp32"] = (out.float() - ref).abs().max().item()
                r["rel_l2_vs_fp32"] = ((out.float() - ref).norm() / ref.norm()).item()
        r.update(burst_ms=tb, burst_tflops=fl / tb / 1e9, sustained_ms=ts, sustained_tflops=fl / ts / 1e9)
        rows.append(r)
    print("ROWS " + json.dumps(rows), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="51")
    ap.add_argument("--child", default=None)
    ap.add_argument("--out", default="gpurun_out/attn_variants.json")
    args = ap.parse_args()
    if args.child is not None:
        child(args.child)
        return
    res = []
    for var in args.variants.split(",") + ["cudnn"]:
        env = dict(os.environ)
        if var != "cudnn":
            env["B2F_ATTN_VARIANT"] = var
        try:
            cp = subprocess.run([sys.executable, __file__, "--child", var], env=env, capture_output=True, text=True,
                                timeout=100)
        except subprocess.TimeoutExpired:
            res.append(dict(variant=var, error="timeout (100 s)"))
            print(f"variant {var}: timeout", flush=True)
            continue
        rows = [json.loads(l[5:]) for l in cp.stdout.splitlines() if l.startswith("ROWS ")]
        if cp.returncode != 0 or not rows:
            res.append(dict(variant=var, error=(cp.stderr or cp.stdout)[-600:]))
            print(f"variant {var}: failed rc={cp.returncode}\n{cp.stderr[-600:]}", flush=True)
            continue
        for r in rows[0]:
            print(json.dumps(r), flush=True)
        res.extend(rows[0])
    Path(args.out).parent.mkdir(exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
