"""Tensor-level wrappers over the C-ABI (raw pointers + shapes from torch tensors).

No arithmetic happens here - each function validates devices/dtypes/strides,
allocates outputs through torch's caching allocator and forwards to one
`xclip_*` entry point on torch's current CUDA stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.XClipB200Error(f"{name} must be a CUDA tensor (x_clip_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.XClipB200Error(f"{name} must be {dtype}, got {t.dtype}")


def _rows2d(t: torch.Tensor, name: str) -> None:
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.XClipB200Error(f"{name} must be 2-D with a contiguous last dim, got "
                                  f"shape {tuple(t.shape)} strides {t.stride()}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_major: int = 0, b_major: int = 0,
         out: Optional[torch.Tensor] = None, out_dtype=BF16, alpha: float = 1.0,
         bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         res_row_mod: int = 0, accumulate: bool = False) -> torch.Tensor:
    """out[M,N] (+)= alpha * A @ B^T (+bias) (+residual).  See xclip_gemm_bf16."""
    _need(a, BF16, "a"); _need(b, BF16, "b")
    _rows2d(a, "a"); _rows2d(b, "b")
    if a_major == 0:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_major == 0:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise _lib.XClipB200Error(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        if accumulate:
            out = torch.zeros((M, N), device=a.device, dtype=F32)
        else:
            out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    _rows2d(out, "out")
    if tuple(out.shape) != (M, N):
        raise _lib.XClipB200Error(f"gemm: out shape {tuple(out.shape)} != {(M, N)}")
    if out.dtype not in (BF16, F32):
        raise _lib.XClipB200Error("gemm: out must be bf16 or f32")
    if bias is not None:
        _need(bias, F32, "bias")
    ldr = 0
    if residual is not None:
        _need(residual, BF16, "residual"); _rows2d(residual, "residual")
        ldr = residual.stride(0)
    _lib.call("xclip_gemm_bf16", a.data_ptr(), a.stride(0), a_major, b.data_ptr(), b.stride(0),
              b_major, out.data_ptr(), out.stride(0), 1 if out.dtype == F32 else 0, M, N, K,
              float(alpha), _ptr(bias), _ptr(residual), ldr, int(res_row_mod),
              1 if accumulate else 0, _stream())
    return out
