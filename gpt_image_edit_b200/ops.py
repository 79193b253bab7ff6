"""Tensor-level wrappers over the C ABI (torch owns storage; libb2f does the work).

Each function validates dtype/device/contiguity on the host, then passes raw device pointers and
the current CUDA stream to libb2f.  No function here computes anything in torch.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

EPI_BIAS, EPI_GELU_TANH, EPI_SILU, EPI_GATE_RESID = 0, 1, 2, 3


def _req(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.B2FError(f"{name}: libb2f runs on CUDA tensors only (got {t.device}); there is no CPU path")
    if t.dtype != torch.bfloat16:
        raise _lib.B2FError(f"{name}: expected bfloat16, got {t.dtype}")
    if t.stride(-1) != 1:
        raise _lib.B2FError(f"{name}: innermost dimension must be contiguous")


def linear(
    x: torch.Tensor,
    weight: torch.Tensor,
    bias: torch.Tensor | None = None,
    *,
    epilogue: int = EPI_BIAS,
    out: torch.Tensor | None = None,
    resid: torch.Tensor | None = None,
    gate: torch.Tensor | None = None,
    rows_per_batch: int = 0,
) -> torch.Tensor:
    """out[M,N] = epilogue(x[M,K] @ weight[N,K]^T + bias) via b2f_gemm_bf16 (tcgen05)."""
    _req(x, "x")
    _req(weight, "weight")
    x2 = x.reshape(-1, x.shape[-1]) if x.dim() != 2 else x
    M, K = x2.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.bfloat16)
    out2 = out.reshape(-1, out.shape[-1]) if out.dim() != 2 else out
    gate_ld = 0
    ldr = 0
    if epilogue == EPI_GATE_RESID:
        _req(resid, "resid")
        _req(gate, "gate")
        resid2 = resid.reshape(-1, resid.shape[-1]) if resid.dim() != 2 else resid
        ldr = resid2.stride(0)
        gate_ld = gate.stride(0) if gate.dim() == 2 else 0
    check(
        _lib.lib.b2f_gemm_bf16(
            ptr(x2), x2.stride(0), ptr(weight), weight.stride(0), ptr(bias), ptr(out2), out2.stride(0),
            M, N, K, epilogue, ptr(resid), ldr, ptr(gate), gate_ld, rows_per_batch, stream_ptr(),
        ),
        "b2f_gemm_bf16",
    )
    return out if x.dim() == 2 else out.reshape(*x.shape[:-1], N)


def attention(
    q: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
    *,
    out: torch.Tensor | None = None,
    causal: bool = False,
    scale: float | None = None,
) -> torch.Tensor:
    """softmax(q k^T * scale) v via b2f_attention_fwd.  q [B,Sq,H,128], k/v [B,Skv,Hkv,128] as
    (possibly strided) views whose last two dims are contiguous; out [B,Sq,H*128]."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, n)
        if t.dim() != 4 or t.stride(2) != t.shape[3] or t.stride(0) != t.shape[1] * t.stride(1):
            raise _lib.B2FError(f"{n}: expected a [B,S,H,dh] view with contiguous heads and batch stride S*ld")
    B, Sq, H, dh = q.shape
    Skv, Hkv = k.shape[1], k.shape[2]
    if out is None:
        out = torch.empty((B, Sq, H * dh), device=q.device, dtype=torch.bfloat16)
    _req(out, "out")
    if scale is None:
        scale = dh ** -0.5
    check(
        _lib.lib.b2f_attention_fwd(
            ptr(q), q.stride(1), ptr(k), k.stride(1), ptr(v), v.stride(1), ptr(out), out.stride(1),
            B, H, Hkv, Sq, Skv, dh, float(scale), int(causal), stream_ptr(),
        ),
        "b2f_attention_fwd",
    )
    return out
