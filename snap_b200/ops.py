"""Thin Python wrappers over the C ABI: torch tensors in, device pointers + current stream out.

torch is plumbing only (device memory, streams); no torch op computes anything on the hot path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.SnapB200Error("libsnapb200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _require(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise _lib.SnapB200Error(f"{name}: must be a CUDA tensor (no CPU fallback)")


def gemm(
    a: torch.Tensor,                # bf16 [a_rows, a_ld-strided], 2D view with unit inner stride
    b: torch.Tensor,                # bf16 [N, num_seg*seg_k]
    out: torch.Tensor,              # bf16 or f32 [rows, ldo-strided]
    *,
    m_rows: Optional[int] = None,
    n: Optional[int] = None,
    seg_off: Sequence[int] = (0,),
    seg_k: Optional[int] = None,
    a_col0: int = 0,
    residual: Optional[torch.Tensor] = None,
    bias: Optional[torch.Tensor] = None,
    row_mask: Optional[torch.Tensor] = None,
    relu: bool = False,
    remap: Optional[Sequence[int]] = None,   # (R, C, r0, c0, Ho, Wo)
    bn: int = 0,
) -> torch.Tensor:
    """out = epilogue(sum_s A[m + seg_off[s], a_col0 : a_col0+seg_k] @ B[:, s*seg_k:(s+1)*seg_k].T)."""
    _require(a, torch.bfloat16, "a")
    _require(b, torch.bfloat16, "b")
    assert a.dim() == 2 and b.dim() == 2 and out.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    p = _lib.GemmParams()
    p.a, p.a_rows, p.a_cols, p.a_ld = _ptr(a), a.shape[0], a.shape[1], a.stride(0)
    p.b, p.b_rows, p.b_cols, p.b_ld = _ptr(b), b.shape[0], b.shape[1], b.stride(0)
    num_seg = len(seg_off)
    p.num_seg = num_seg
    p.seg_k = seg_k if seg_k is not None else b.shape[1] // num_seg
    p.m_rows = m_rows if m_rows is not None else a.shape[0]
    p.n = n if n is not None else b.shape[0]
    p.a_col0 = a_col0
    for i, o in enumerate(seg_off):
        p.seg_off[i] = int(o)
    p.out, p.ldo = _ptr(out), out.stride(0)
    if out.dtype == torch.float32:
        p.out_f32 = 1
    elif out.dtype == torch.bfloat16:
        p.out_f32 = 0
    else:
        raise TypeError("out must be bf16 or f32")
    if residual is not None:
        _require(residual, torch.bfloat16, "residual")
        p.residual, p.ldr = _ptr(residual), residual.stride(0)
    if bias is not None:
        _require(bias, torch.float32, "bias")
        p.bias = _ptr(bias)
    if row_mask is not None:
        _require(row_mask, torch.uint8, "row_mask")
        p.row_mask = _ptr(row_mask)
    p.relu = int(relu)
    if remap is not None:
        p.remap = 1
        p.rm_R, p.rm_C, p.rm_r0, p.rm_c0, p.rm_Ho, p.rm_Wo = (int(x) for x in remap)
    p.bn = bn
    _lib.check(_lib.lib().snapb200_gemm_bf16(C.byref(p), _stream()))
    return out
