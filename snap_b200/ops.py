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
    gn_acc: Optional[torch.Tensor] = None,       # f64 [GN_REPLICAS, n_img, 32, 2] accumulators (zeroed by the caller)
    gn_acc_relu: Optional[torch.Tensor] = None,
    gn_rows_per_img: int = 0,
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
    if gn_acc is not None:
        _require(gn_acc, torch.float64, "gn_acc")
        assert gn_acc.dim() == 4 and gn_acc.shape[0] == GN_REPLICAS and gn_acc.is_contiguous()
        p.gn_acc, p.gn_rows_per_img, p.gn_replica_stride = _ptr(gn_acc), gn_rows_per_img, gn_acc.stride(0)
        if gn_acc_relu is not None:
            _require(gn_acc_relu, torch.float64, "gn_acc_relu")
            p.gn_acc_relu = _ptr(gn_acc_relu)
    _lib.check(_lib.lib().snapb200_gemm_bf16(C.byref(p), _stream()))
    return out


def conv_gn(x: torch.Tensor, n_img: int, H: int, W: int, Cc: int, acc: torch.Tensor, scale: torch.Tensor,
            bias: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, taps: int = 1, stride: int = 1,
            pre_relu: bool = False, post_relu: bool = True, residual: Optional[torch.Tensor] = None,
            gn_acc: Optional[torch.Tensor] = None, gn_acc_relu: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = conv(act(GroupNorm(pre(x)))) in one launch; x is the RAW activation [n_img*H*W, Cc] bf16, its
    GroupNorm statistics are in `acc` (f64 [GN_REPLICAS, n_img, 32, 2])."""
    _require(x, torch.bfloat16, "x")
    _require(b, torch.bfloat16, "b")
    _require(out, torch.bfloat16, "out")
    _require(acc, torch.float64, "acc")
    p = _lib.ConvGnParams()
    p.x, p.n_img, p.H, p.W, p.C = _ptr(x), n_img, H, W, Cc
    p.acc, p.replica_stride = _ptr(acc), acc.stride(0)
    p.scale, p.bias = _ptr(scale), _ptr(bias)
    p.pre_relu, p.post_relu, p.taps, p.stride = int(pre_relu), int(post_relu), taps, stride
    p.b, p.b_rows, p.b_cols, p.b_ld = _ptr(b), b.shape[0], b.shape[1], b.stride(0)
    p.n = b.shape[0]
    p.out, p.ldo = _ptr(out), out.stride(0)
    if residual is not None:
        _require(residual, torch.bfloat16, "residual")
        p.residual, p.ldr = _ptr(residual), residual.stride(0)
    if gn_acc is not None:
        _require(gn_acc, torch.float64, "gn_acc")
        p.gn_acc, p.gn_replica_stride = _ptr(gn_acc), gn_acc.stride(0)
        if gn_acc_relu is not None:
            p.gn_acc_relu = _ptr(gn_acc_relu)
    _lib.check(_lib.lib().snapb200_conv_gn_bf16(C.byref(p), _stream()))
    return out


def conv3x3_halo_supported(Cc: int, n: int, W: int) -> bool:
    return bool(_lib.lib().snapb200_conv3x3_halo_supported(int(Cc), int(n), int(W)))


def conv3x3_halo(a: torch.Tensor, n_img: int, H: int, W: int, Cc: int, b: torch.Tensor, out: torch.Tensor,
                 gn_acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 / stride-1 conv of the zero-bordered input a bf16 [n_img*(H+2)*(W+2) (+ slack), Cc] with b bf16 [n, 9*Cc] ->
    out bf16 [n_img*H*W, n] (dense); gn_acc: f64 [GN_REPLICAS, n_img, 32, 2] statistics of the output."""
    _require(a, torch.bfloat16, "a")
    _require(b, torch.bfloat16, "b")
    _require(out, torch.bfloat16, "out")
    p = _lib.Conv3x3Params()
    p.a, p.n_img, p.H, p.W, p.C = _ptr(a), n_img, H, W, Cc
    p.b, p.b_ld, p.n = _ptr(b), b.stride(0), b.shape[0]
    p.out, p.ldo = _ptr(out), out.stride(0)
    if gn_acc is not None:
        _require(gn_acc, torch.float64, "gn_acc")
        p.gn_acc, p.gn_replica_stride = _ptr(gn_acc), gn_acc.stride(0)
    _lib.check(_lib.lib().snapb200_conv3x3_halo_bf16(C.byref(p), _stream()))
    return out


def selftest_shifted_desc(a: torch.Tensor, b: torch.Tensor, shift: int, mode: int) -> torch.Tensor:
    """a bf16 [256,64], b bf16 [64,64] -> f32 [128,64] = a[shift:shift+128] @ b^T through a row-shifted smem descriptor."""
    _require(a, torch.bfloat16, "a")
    _require(b, torch.bfloat16, "b")
    assert tuple(a.shape) == (256, 64) and tuple(b.shape) == (64, 64) and a.is_contiguous() and b.is_contiguous()
    out = torch.zeros((128, 64), dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().snapb200_selftest_shifted_desc(C.c_void_p(_ptr(a)), C.c_void_p(_ptr(b)), int(shift), int(mode),
                                                         C.c_void_p(_ptr(out)), _stream()))
    return out


# --------------------------------------------------------------------------------------------
# image-encoder kernels
# --------------------------------------------------------------------------------------------
def std_weights_batched(descs: torch.Tensor, mapA: Optional[torch.Tensor], mapB: torch.Tensor) -> None:
    nA = 0 if mapA is None else mapA.shape[0]
    _lib.check(_lib.lib().snapb200_std_weights_batched(
        C.c_void_p(_ptr(descs)), C.c_void_p(_ptr(mapA) if nA else None), C.c_int(nA),
        C.c_void_p(_ptr(mapB)), C.c_int(mapB.shape[0]), _stream()))


def root_im2col(images: torch.Tensor, Hp: int, Wp: int, KH: int, KW: int, stride: int, pad: int,
                out: torch.Tensor) -> None:
    _require(images, torch.float32, "images")
    n, H, W, c = images.shape
    assert c == 3 and images.is_contiguous()
    _lib.check(_lib.lib().snapb200_root_im2col(
        C.c_void_p(_ptr(images)), n, H, W, Hp, Wp, KH, KW, stride, pad, C.c_void_p(_ptr(out)),
        C.c_int(out.shape[1]), _stream()))


def root_packed_geometry(Hp: int, Wp: int, KH: int, KW: int, stride: int, pad: int):
    """(cp, Hq, Wq, Ho, Wo) of the packed image of the implicit root conv (see include/snapb200.h)."""
    cp = 8 // stride  # window step stride*cp*2 B must be 16 B
    assert stride in (1, 2) and KW * cp <= 32
    Ho, Wo = (Hp + 2 * pad - KH) // stride + 1, (Wp + 2 * pad - KW) // stride + 1
    Hq = max(Hp + 2 * pad, (Ho - 1) * stride + KH)
    Wq = max(Wp + 2 * pad, (Wo - 1) * stride + 32 // cp)
    Wq += Wq % 2
    return cp, Hq, Wq, Ho, Wo


def root_pack_image(images: torch.Tensor, Hp: int, Wp: int, pad: int, cp: int, Hq: int, Wq: int,
                    out: torch.Tensor) -> None:
    _require(images, torch.float32, "images")
    _require(out, torch.bfloat16, "out")
    n, H, W, c = images.shape
    assert c == 3 and images.is_contiguous() and out.numel() >= n * Hq * Wq * cp
    _lib.check(_lib.lib().snapb200_root_pack_image(C.c_void_p(_ptr(images)), n, H, W, Hp, Wp, pad, cp, Hq, Wq,
                                                   C.c_void_p(_ptr(out)), _stream()))


def root_pack_weights(b_std: torch.Tensor, cout: int, KH: int, KW: int, cp: int, out: torch.Tensor) -> None:
    _require(b_std, torch.bfloat16, "b_std")
    assert out.shape[1] == KH * 32 and out.shape[0] >= cout
    _lib.check(_lib.lib().snapb200_root_pack_weights(C.c_void_p(_ptr(b_std)), cout, b_std.stride(0), KH, KW, cp,
                                                     C.c_void_p(_ptr(out)), _stream()))


def root_conv(packed: torch.Tensor, n: int, Hq: int, Wq: int, cp: int, KH: int, stride: int, Ho: int, Wo: int,
              b: torch.Tensor, cout: int, out: torch.Tensor, gn_acc: Optional[torch.Tensor] = None) -> None:
    p = _lib.RootConvParams()
    p.packed, p.n_img, p.Hq, p.Wq, p.cp = _ptr(packed), n, Hq, Wq, cp
    p.KH, p.stride, p.Ho, p.Wo = KH, stride, Ho, Wo
    p.b, p.n, p.out, p.ldo = _ptr(b), cout, _ptr(out), out.stride(0)
    if gn_acc is not None:
        _require(gn_acc, torch.float64, "gn_acc")
        p.gn_acc, p.gn_replica_stride = _ptr(gn_acc), gn_acc.stride(0)
    _lib.check(_lib.lib().snapb200_root_conv_bf16(C.byref(p), _stream()))


def maxpool3x3s2(x: torch.Tensor, n: int, H: int, W: int, Cc: int, y: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_maxpool3x3s2(C.c_void_p(_ptr(x)), n, H, W, Cc, C.c_void_p(_ptr(y)), _stream()))


GN_REPLICAS = 8


def gn_stats(x: torch.Tensor, n: int, hw: int, Cc: int, pre_relu: bool, acc: torch.Tensor) -> None:
    """Accumulate raw GroupNorm statistics of x into replica 0 of acc f64 [GN_REPLICAS, n, 32, 2]
    (must be zeroed by the caller)."""
    _require(acc, torch.float64, "acc")
    assert acc.dim() == 4 and acc.shape[0] == GN_REPLICAS
    _lib.check(_lib.lib().snapb200_gn_stats(C.c_void_p(_ptr(x)), n, hw, Cc, int(pre_relu),
                                            C.c_void_p(_ptr(acc)), _stream()))


LAYOUT_DENSE, LAYOUT_PADDED, LAYOUT_PHASE = 0, 1, 2


def gn_apply(x: torch.Tensor, n: int, H: int, W: int, Cc: int, acc: torch.Tensor, scale: torch.Tensor,
             bias: torch.Tensor, pre_relu: bool, post_relu: bool, layout: int, out: torch.Tensor,
             out_sub: Optional[torch.Tensor] = None) -> None:
    _lib.check(_lib.lib().snapb200_gn_apply(
        C.c_void_p(_ptr(x)), n, H, W, Cc, C.c_void_p(_ptr(acc)), C.c_int(acc.stride(0)), C.c_void_p(_ptr(scale)),
        C.c_void_p(_ptr(bias)), int(pre_relu), int(post_relu), layout, C.c_void_p(_ptr(out)),
        C.c_void_p(_ptr(out_sub)), _stream()))


def upsample2x(x: torch.Tensor, n: int, h: int, w: int, Cc: int, y: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_upsample2x(C.c_void_p(_ptr(x)), n, h, w, Cc, C.c_void_p(_ptr(y)), _stream()))


def crop_relu(x: torch.Tensor, n: int, Hs: int, Ws: int, Cc: int, h: int, w: int, relu: bool,
              y: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_crop_relu(C.c_void_p(_ptr(x)), n, Hs, Ws, Cc, h, w, int(relu),
                                             C.c_void_p(_ptr(y)), _stream()))


# --------------------------------------------------------------------------------------------
# lift
# --------------------------------------------------------------------------------------------
def lift_gather_pool(p: "_lib.LiftParams", views: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                     zs: torch.Tensor, stats: torch.Tensor, valid: torch.Tensor,
                     dbg_vis: Optional[torch.Tensor] = None, dbg_taps: Optional[torch.Tensor] = None) -> None:
    _require(fimg, torch.bfloat16, "fimg")
    _lib.check(_lib.lib().snapb200_lift_gather_pool(
        C.byref(p), C.c_void_p(_ptr(views)), C.c_void_p(_ptr(fimg)), C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)),
        C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(stats)), C.c_void_p(_ptr(valid)),
        C.c_void_p(_ptr(dbg_vis)), C.c_void_p(_ptr(dbg_taps)), _stream()))


def lift_select_pool(p: "_lib.LiftParams", top_k: int, max_view_distance: Optional[float], views: torch.Tensor,
                     view_centers: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                     zs: torch.Tensor, stats: torch.Tensor, valid: torch.Tensor,
                     dbg_idx: Optional[torch.Tensor] = None, dbg_vis: Optional[torch.Tensor] = None,
                     dbg_taps: Optional[torch.Tensor] = None) -> None:
    """V > top_k path: view selection + selective sampling + pooling (`streetview_encoder.py:127-138,80-105`)."""
    _require(fimg, torch.bfloat16, "fimg")
    _require(view_centers, torch.float32, "view_centers")
    _lib.check(_lib.lib().snapb200_lift_select_pool(
        C.byref(p), C.c_int(top_k), C.c_float(-1.0 if max_view_distance is None else max_view_distance),
        C.c_void_p(_ptr(views)), C.c_void_p(_ptr(view_centers)), C.c_void_p(_ptr(fimg)), C.c_void_p(_ptr(xs)),
        C.c_void_p(_ptr(ys)), C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(stats)), C.c_void_p(_ptr(valid)),
        C.c_void_p(_ptr(dbg_idx)), C.c_void_p(_ptr(dbg_vis)), C.c_void_p(_ptr(dbg_taps)), _stream()))


def lift_fused(p: "_lib.LiftParams", views: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
               zs: torch.Tensor, w1t: torch.Tensor, w256: torch.Tensor, b1: torch.Tensor, w2t: torch.Tensor,
               b2: torch.Tensor, plane: torch.Tensor, pvalid: torch.Tensor, counter: torch.Tensor,
               scratch: torch.Tensor) -> None:
    _require(fimg, torch.bfloat16, "fimg")
    _require(w1t, torch.bfloat16, "w1t")
    _require(w2t, torch.bfloat16, "w2t")
    for t, n in ((w256, "w256"), (b1, "b1"), (b2, "b2")):
        _require(t, torch.float32, n)
    _require(counter, torch.int32, "counter")
    assert w1t.shape[0] >= 256 and w2t.shape == (128, 256) and w2t.is_contiguous()
    _lib.check(_lib.lib().snapb200_lift_fused(
        C.byref(p), C.c_void_p(_ptr(views)), C.c_void_p(_ptr(fimg)), C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)),
        C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(w1t)), C.c_longlong(w1t.stride(0)), C.c_void_p(_ptr(w256)),
        C.c_void_p(_ptr(b1)), C.c_void_p(_ptr(w2t)), C.c_void_p(_ptr(b2)), C.c_void_p(_ptr(plane)),
        C.c_void_p(_ptr(pvalid)), C.c_void_p(_ptr(counter)), C.c_void_p(_ptr(scratch)),
        C.c_size_t(scratch.numel() * scratch.element_size()), _stream()))


def lift_fused_scratch_bytes() -> int:
    f = _lib.lib().snapb200_lift_fused_scratch_bytes
    f.restype = C.c_size_t
    return int(f())


def lift_fused_batched(p: "_lib.LiftParams", B: int, views: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor,
                       ys: torch.Tensor, zs: torch.Tensor, w1t: torch.Tensor, w256: torch.Tensor, b1: torch.Tensor,
                       w2t: torch.Tensor, b2: torch.Tensor, plane: torch.Tensor, pvalid: torch.Tensor,
                       counter: torch.Tensor, scratch: torch.Tensor) -> None:
    """All B scenes in one launch of the warp-specialised lift: views [B, >= V * words] i32 (SnapLiftView tables),
    fimg bf16 [B, rows, CF], zs f32 [B, Z], plane bf16 [B, X*Y, 128], pvalid u8 [B, X*Y]."""
    _require(fimg, torch.bfloat16, "fimg")
    _require(w1t, torch.bfloat16, "w1t")
    _require(w2t, torch.bfloat16, "w2t")
    for t, n in ((w256, "w256"), (b1, "b1"), (b2, "b2"), (zs, "zs")):
        _require(t, torch.float32, n)
    _require(counter, torch.int32, "counter")
    assert w1t.shape[0] >= 256 and w2t.shape == (128, 256) and w2t.is_contiguous()
    assert views.dim() == 2 and fimg.dim() == 3 and zs.dim() == 2 and views.shape[0] >= B and fimg.shape[0] >= B
    assert plane.is_contiguous() and pvalid.is_contiguous()
    view_words = C.sizeof(_lib.LiftView) // 4
    assert views.stride(0) % view_words == 0 and views.stride(1) == 1
    _lib.check(_lib.lib().snapb200_lift_fused_batched(
        C.byref(p), C.c_int(B), C.c_void_p(_ptr(views)), C.c_longlong(views.stride(0) // view_words),
        C.c_void_p(_ptr(fimg)), C.c_longlong(fimg.stride(0)), C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)),
        C.c_void_p(_ptr(zs)), C.c_longlong(zs.stride(0)), C.c_void_p(_ptr(w1t)), C.c_longlong(w1t.stride(0)),
        C.c_void_p(_ptr(w256)), C.c_void_p(_ptr(b1)), C.c_void_p(_ptr(w2t)), C.c_void_p(_ptr(b2)),
        C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(pvalid)), C.c_void_p(_ptr(counter)), C.c_void_p(_ptr(scratch)),
        C.c_size_t(scratch.numel() * scratch.element_size()), _stream()))


def lift_fused_batched_scratch_bytes() -> int:
    f = _lib.lib().snapb200_lift_fused_batched_scratch_bytes
    f.restype = C.c_size_t
    return int(f())


def lift_observe(p: "_lib.LiftParams", views: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor,
                 zs: torch.Tensor, obs: torch.Tensor, vis: torch.Tensor) -> None:
    """Per-observation rows [f(128) | log10 depth | ray(3) | 0...] of the depth_mlp branch (`streetview_encoder.py:262-267`)."""
    _require(fimg, torch.bfloat16, "fimg")
    _require(obs, torch.bfloat16, "obs")
    assert obs.shape[1] == 160 and obs.is_contiguous() and vis.dtype == torch.uint8
    _lib.check(_lib.lib().snapb200_lift_observe(
        C.byref(p), C.c_void_p(_ptr(views)), C.c_void_p(_ptr(fimg)), C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)),
        C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(obs)), C.c_void_p(_ptr(vis)), _stream()))


def lift_pool_observations(V: int, N: int, obs: torch.Tensor, d: torch.Tensor, vis: torch.Tensor, stats: torch.Tensor,
                           valid: torch.Tensor) -> None:
    """f' = bf16(f + d), plain mean / variance over the visible views -> statistics rows [mean | var | 0...], valid."""
    _require(d, torch.bfloat16, "d")
    assert d.shape[1] == 128 and d.is_contiguous() and stats.is_contiguous()
    _lib.check(_lib.lib().snapb200_lift_pool_observations(
        C.c_int(V), C.c_longlong(N), C.c_void_p(_ptr(obs)), C.c_void_p(_ptr(d)), C.c_void_p(_ptr(vis)),
        C.c_int(stats.shape[1]), C.c_void_p(_ptr(stats)), C.c_void_p(_ptr(valid)), _stream()))


def vertical_max(volume: torch.Tensor, valid: torch.Tensor, cells: int, Z: int, Cc: int,
                 plane: torch.Tensor, pvalid: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_vertical_max(
        C.c_void_p(_ptr(volume)), C.c_void_p(_ptr(valid)), C.c_longlong(cells), Z, Cc,
        C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(pvalid)), _stream()))


POOL_MODES = {"sum": 1, "mean": 2, "softmax": 3, "weighted": 4}


def vertical_pool(mode: str, volume: torch.Tensor, valid: torch.Tensor, cells: int, Z: int, Cc: int,
                  conf_w: Optional[torch.Tensor], conf_b: float, plane: torch.Tensor, pvalid: torch.Tensor,
                  scores: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None) -> None:
    """VerticalPooling 'sum' / 'mean' / 'softmax' / 'weighted' (`bev_mapper.py:56-88`)."""
    _require(volume, torch.bfloat16, "volume")
    if conf_w is not None:
        _require(conf_w, torch.float32, "conf_w")
    _lib.check(_lib.lib().snapb200_vertical_pool(
        C.c_int(POOL_MODES[mode]), C.c_void_p(_ptr(volume)), C.c_void_p(_ptr(valid)), C.c_longlong(cells), Z, Cc,
        C.c_void_p(_ptr(conf_w)), C.c_float(conf_b), C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(pvalid)),
        C.c_void_p(_ptr(scores)), C.c_void_p(_ptr(weights)), _stream()))


def confidence(plane: torch.Tensor, valid: torch.Tensor, cells: int, Cc: int, conf_w: torch.Tensor, conf_b: float,
               out: torch.Tensor) -> None:
    """`bev_mapper.py:292-295`: where(valid, log_sigmoid(Dense(C -> 1)(plane)), 0) -> f32 [cells]."""
    _require(plane, torch.bfloat16, "plane")
    _require(conf_w, torch.float32, "conf_w")
    _require(out, torch.float32, "out")
    _lib.check(_lib.lib().snapb200_confidence(C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(valid)), C.c_longlong(cells), Cc,
                                              C.c_void_p(_ptr(conf_w)), C.c_float(conf_b), C.c_void_p(_ptr(out)), _stream()))


def mask_rows(x: torch.Tensor, valid: torch.Tensor, rows: int, Cc: int, y: torch.Tensor) -> None:
    _require(x, torch.bfloat16, "x")
    _lib.check(_lib.lib().snapb200_mask_rows(C.c_void_p(_ptr(x)), C.c_void_p(_ptr(valid)), C.c_longlong(rows), Cc,
                                             C.c_void_p(_ptr(y)), _stream()))


def valid_any(valid: torch.Tensor, cells: int, Z: int, out: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_valid_any(C.c_void_p(_ptr(valid)), C.c_longlong(cells), Z,
                                             C.c_void_p(_ptr(out)), _stream()))


def match_head(plane: torch.Tensor, valid: torch.Tensor, cells: int, Cc: int, kernel: torch.Tensor,
               bias: torch.Tensor, out: torch.Tensor, normalize: bool = True) -> None:
    """Dense(C -> matching_dim) + optional L2 normalisation + mask (`bev_mapper.py:284-291`)."""
    _require(kernel, torch.float32, "kernel")
    assert kernel.is_contiguous() and out.is_contiguous() and out.shape[-1] == kernel.shape[1]
    _lib.check(_lib.lib().snapb200_match_head_ex(
        C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(valid)), C.c_longlong(cells), Cc,
        C.c_void_p(_ptr(kernel)), C.c_void_p(_ptr(bias)), C.c_int(kernel.shape[1]), C.c_int(int(normalize)),
        C.c_void_p(_ptr(out)), _stream()))


def fuse_max(a: torch.Tensor, va: torch.Tensor, b: torch.Tensor, vb: Optional[torch.Tensor], cells: int,
             Cc: int, out: torch.Tensor, vout: torch.Tensor) -> None:
    _lib.check(_lib.lib().snapb200_fuse_max(
        C.c_void_p(_ptr(a)), C.c_void_p(_ptr(va)), C.c_void_p(_ptr(b)), C.c_void_p(_ptr(vb)),
        C.c_longlong(cells), Cc, C.c_void_p(_ptr(out)), C.c_void_p(_ptr(vout)), _stream()))


# --------------------------------------------------------------------------------------------
# exhaustive pose voting
# --------------------------------------------------------------------------------------------
def xcorr_padded_cols(G: int) -> int:
    return int(_lib.lib().snapb200_xcorr_padded_cols(C.c_int(G)))


def xcorr_padded_rotations(R: int) -> int:
    return int(_lib.lib().snapb200_xcorr_padded_rotations(C.c_int(R)))


def rot_templates(feats: torch.Tensor, valid: torch.Tensor, conf: Optional[torch.Tensor], rot_host,
                  centers: torch.Tensor, cell_size: float, R: int, templates: torch.Tensor,
                  t_valid: torch.Tensor) -> None:
    _require(feats, torch.bfloat16, "feats")
    B, G, _, D = feats.shape
    rot = (C.c_float * (R // 4 * 4))(*[float(x) for x in rot_host.reshape(-1)])
    _lib.check(_lib.lib().snapb200_rot_templates(
        C.c_void_p(_ptr(feats)), C.c_void_p(_ptr(valid)), C.c_void_p(_ptr(conf)), rot,
        C.c_void_p(_ptr(centers)), C.c_float(cell_size), B, R, G, D, C.c_void_p(_ptr(templates)),
        C.c_void_p(_ptr(t_valid)), _stream()))


def xcorr_pad_map(m: torch.Tensor, out: torch.Tensor) -> None:
    _require(m, torch.bfloat16, "m")
    B, G, _, D = m.shape
    _lib.check(_lib.lib().snapb200_xcorr_pad_map(C.c_void_p(_ptr(m)), B, G, D, C.c_void_p(_ptr(out)), _stream()))


def xcorr_count(t_valid: torch.Tensor, m_valid: torch.Tensor, cnt: torch.Tensor, den: torch.Tensor) -> None:
    B, R, G, _ = t_valid.shape
    _lib.check(_lib.lib().snapb200_xcorr_count(
        C.c_void_p(_ptr(t_valid)), C.c_void_p(_ptr(m_valid)), B, R, G, C.c_void_p(_ptr(cnt)),
        C.c_void_p(_ptr(den)), _stream()))


def xcorr_scores(templates: torch.Tensor, m_pad: torch.Tensor, cnt: Optional[torch.Tensor],
                 den: Optional[torch.Tensor], thr: float, scores: torch.Tensor) -> None:
    """templates: cell-major bf16 [B, G, G, padded_rotations(R), D]; scores f32 [B, R, 2G-1, 2G-1]."""
    B, G, _, RP, D = templates.shape
    R = scores.shape[1]
    assert RP == xcorr_padded_rotations(R)
    _require(scores, torch.float32, "scores")
    _lib.check(_lib.lib().snapb200_xcorr_scores(
        C.c_void_p(_ptr(templates)), C.c_void_p(_ptr(m_pad)), C.c_void_p(_ptr(cnt)),
        C.c_void_p(_ptr(den)), B, R, G, D, C.c_float(thr), C.c_void_p(_ptr(scores)), _stream()))


def xcorr_sw_supported(R: int, G: int) -> bool:
    return xcorr_padded_rotations(R) == 48 and G % 4 == 0 and G <= 129


def xcorr_scores_sw(templates: torch.Tensor, m_pad: torch.Tensor, cnt: Optional[torch.Tensor],
                    den: Optional[torch.Tensor], thr: float, scores: torch.Tensor) -> None:
    """Sliding-window variant of `xcorr_scores` (same arguments)."""
    B, G, _, RP, D = templates.shape
    R = scores.shape[1]
    _require(scores, torch.float32, "scores")
    _lib.check(_lib.lib().snapb200_xcorr_scores_sw(
        C.c_void_p(_ptr(templates)), C.c_void_p(_ptr(m_pad)), C.c_void_p(_ptr(cnt)),
        C.c_void_p(_ptr(den)), B, R, G, D, C.c_float(thr), C.c_void_p(_ptr(scores)), _stream()))


def xcorr_rows_supported(R: int, G: int) -> bool:
    return R % 4 == 0 and R <= 64 and G % 2 == 0 and 8 <= G <= 129


def xcorr_rows_workspace_bytes(B: int, R: int, G: int, D: int) -> int:
    f = _lib.lib().snapb200_xcorr_scores_rows_workspace
    f.restype = C.c_size_t
    return int(f(B, R, G, D))


def xcorr_scores_rows(templates: torch.Tensor, m_pad: torch.Tensor, cnt: Optional[torch.Tensor],
                      den: Optional[torch.Tensor], thr: float, scores: torch.Tensor,
                      workspace: Optional[torch.Tensor] = None) -> None:
    """Map-row-major variant of `xcorr_scores` (same arguments + workspace of `xcorr_rows_workspace_bytes`)."""
    B, G, _, RP, D = templates.shape
    R = scores.shape[1]
    _require(scores, torch.float32, "scores")
    need = xcorr_rows_workspace_bytes(B, R, G, D)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=templates.device)
    assert workspace.numel() * workspace.element_size() >= need
    _lib.check(_lib.lib().snapb200_xcorr_scores_rows(
        C.c_void_p(_ptr(templates)), C.c_void_p(_ptr(m_pad)), C.c_void_p(_ptr(cnt)), C.c_void_p(_ptr(den)),
        B, R, G, D, C.c_float(thr), C.c_void_p(_ptr(scores)), C.c_void_p(_ptr(workspace)),
        C.c_size_t(workspace.numel() * workspace.element_size()), _stream()))


# --------------------------------------------------------------------------------------------
# sampling localizer (bev_localizer.py:156-218, pose_estimation.py)
# --------------------------------------------------------------------------------------------
def loc_softmax_stats(sim: torch.Tensor, H: int, W: int, scale: float, row_max: torch.Tensor,
                      chunk_sum: torch.Tensor, row_sum: Optional[torch.Tensor] = None) -> None:
    """sim bf16 [..., H*W] (contiguous) -> row_max f32 [rows], chunk_sum f32 [rows, H], row_sum f32 [rows]."""
    _require(sim, torch.bfloat16, "sim")
    _require(row_max, torch.float32, "row_max")
    _require(chunk_sum, torch.float32, "chunk_sum")
    rows = sim.numel() // (H * W)
    assert sim.is_contiguous() and row_max.numel() == rows and chunk_sum.numel() == rows * H
    _lib.check(_lib.lib().snapb200_loc_softmax_stats(
        C.c_void_p(_ptr(sim)), C.c_longlong(rows), H, W, C.c_float(scale), C.c_void_p(_ptr(row_max)),
        C.c_void_p(_ptr(row_sum)), C.c_void_p(_ptr(chunk_sum)), _stream()))


def loc_point_weights(valid_points: torch.Tensor, conf: Optional[torch.Tensor], exp_t: float,
                      point_scale: torch.Tensor, row_cdf: torch.Tensor) -> None:
    _require(valid_points, torch.uint8, "valid_points")
    B, N = valid_points.shape
    if conf is not None:
        _require(conf, torch.float32, "conf")
        assert conf.shape == (B, N) and conf.is_contiguous()
    assert valid_points.is_contiguous() and point_scale.shape == (B, N) and row_cdf.shape == (B, N)
    _lib.check(_lib.lib().snapb200_loc_point_weights(
        C.c_void_p(_ptr(valid_points)), C.c_void_p(_ptr(conf)), B, N, C.c_float(exp_t),
        C.c_void_p(_ptr(point_scale)), C.c_void_p(_ptr(row_cdf)), _stream()))


def loc_sample(sim: torch.Tensor, row_max: torch.Tensor, chunk_sum: torch.Tensor, row_cdf: torch.Tensor,
               uniforms: torch.Tensor, H: int, W: int, scale: float, indices: torch.Tensor) -> None:
    """uniforms f32 [B,K,2] -> indices i32 [B,K,3] = (point, map row, map column)."""
    _require(sim, torch.bfloat16, "sim")
    _require(uniforms, torch.float32, "uniforms")
    _require(indices, torch.int32, "indices")
    B, N = row_cdf.shape
    K = uniforms.shape[1]
    assert uniforms.shape == (B, K, 2) and indices.shape == (B, K, 3) and uniforms.is_contiguous()
    _lib.check(_lib.lib().snapb200_loc_sample(
        C.c_void_p(_ptr(sim)), C.c_void_p(_ptr(row_max)), C.c_void_p(_ptr(chunk_sum)), C.c_void_p(_ptr(row_cdf)),
        C.c_void_p(_ptr(uniforms)), B, N, H, W, K, C.c_float(scale), C.c_void_p(_ptr(indices)), _stream()))


def loc_ransac_poses(indices: torch.Tensor, i_xy: torch.Tensor, num_poses: int, num_retries: int,
                     cell_size: float, poses: torch.Tensor) -> None:
    """indices i32 [B, num_poses*num_retries*2, 3], i_xy f32 [N,2] or [B,N,2] -> poses f32 [B,num_poses,3]."""
    _require(indices, torch.int32, "indices")
    _require(i_xy, torch.float32, "i_xy")
    _require(poses, torch.float32, "poses")
    B = indices.shape[0]
    assert indices.shape[1] == num_poses * num_retries * 2 and indices.is_contiguous() and i_xy.is_contiguous()
    assert poses.shape == (B, num_poses, 3)
    _lib.check(_lib.lib().snapb200_loc_ransac_poses(
        C.c_void_p(_ptr(indices)), C.c_void_p(_ptr(i_xy)), int(i_xy.dim() == 3), B, i_xy.shape[-2], num_poses,
        num_retries, C.c_float(cell_size), C.c_void_p(_ptr(poses)), _stream()))


def loc_refine_poses(init: torch.Tensor, rot_rad: torch.Tensor, off_x: torch.Tensor, off_y: torch.Tensor,
                     poses: torch.Tensor) -> None:
    _require(init, torch.float32, "init")
    B = init.shape[0]
    nr, nx, ny = rot_rad.numel(), off_x.numel(), off_y.numel()
    assert poses.shape == (B, nr * nx * ny, 3) and init.is_contiguous()
    _lib.check(_lib.lib().snapb200_loc_refine_poses(
        C.c_void_p(_ptr(init)), B, C.c_void_p(_ptr(rot_rad)), nr, C.c_void_p(_ptr(off_x)), nx,
        C.c_void_p(_ptr(off_y)), ny, C.c_void_p(_ptr(poses)), _stream()))


def loc_pose_scoring(sim: torch.Tensor, point_scale: torch.Tensor, i_xy: torch.Tensor,
                     valid_j: Optional[torch.Tensor], poses: torch.Tensor, H: int, W: int, cell_size: float,
                     mask_out_of_bounds: bool, scores: torch.Tensor,
                     workspace: Optional[torch.Tensor] = None) -> None:
    """sim bf16 [B,N,H*W], point_scale f32 [B,N], poses f32 [B,P,3] -> scores f32 [B,P]."""
    _require(sim, torch.bfloat16, "sim")
    _require(poses, torch.float32, "poses")
    _require(scores, torch.float32, "scores")
    _require(i_xy, torch.float32, "i_xy")
    B, N = point_scale.shape
    P = poses.shape[1]
    assert poses.shape == (B, P, 3) and scores.shape == (B, P) and poses.is_contiguous() and sim.is_contiguous()
    assert sim.numel() == B * N * H * W and i_xy.is_contiguous()
    p = _lib.LocScoreParams()
    p.B, p.N, p.H, p.W, p.P = B, N, H, W, P
    p.cell_size = cell_size
    p.mask_out_of_bounds = int(mask_out_of_bounds)
    p.i_xy_batched = int(i_xy.dim() == 3)
    f = _lib.lib().snapb200_loc_pose_scoring_workspace
    f.restype = C.c_size_t
    need = int(f(C.byref(p)))
    if need == 0:
        raise _lib.SnapB200Error(f"loc_pose_scoring: {_lib.lib().snapb200_last_error().decode()}")
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=sim.device)
    assert workspace.numel() * workspace.element_size() >= need
    if valid_j is not None:
        _require(valid_j, torch.uint8, "valid_j")
    _lib.check(_lib.lib().snapb200_loc_pose_scoring(
        C.byref(p), C.c_void_p(_ptr(sim)), C.c_void_p(_ptr(point_scale)), C.c_void_p(_ptr(i_xy)),
        C.c_void_p(_ptr(valid_j)), C.c_void_p(_ptr(poses)), C.c_void_p(_ptr(workspace)),
        C.c_size_t(workspace.numel() * workspace.element_size()), C.c_void_p(_ptr(scores)), _stream()))


def argmax_rows(x: torch.Tensor, start: int, idx: torch.Tensor, rows3: Optional[torch.Tensor] = None,
                best_row3: Optional[torch.Tensor] = None) -> None:
    _require(x, torch.float32, "x")
    _require(idx, torch.int32, "idx")
    rows, cols = x.shape
    assert x.is_contiguous() and idx.numel() == rows
    if rows3 is not None:
        assert rows3.shape == (rows, cols, 3) and rows3.is_contiguous() and best_row3.shape == (rows, 3)
    _lib.check(_lib.lib().snapb200_argmax_rows(
        C.c_void_p(_ptr(x)), rows, cols, start, C.c_void_p(_ptr(idx)), C.c_void_p(_ptr(rows3)),
        C.c_void_p(_ptr(best_row3)), _stream()))


def loc_nll(scores: torch.Tensor, samples: torch.Tensor, best: torch.Tensor, gt: torch.Tensor,
            remove: Optional[Sequence[float]], out: torch.Tensor, dr_samples: Optional[torch.Tensor] = None,
            dt_samples: Optional[torch.Tensor] = None) -> None:
    for t, nm in ((scores, "scores"), (samples, "samples"), (best, "best"), (gt, "gt"), (out, "out")):
        _require(t, torch.float32, nm)
        assert t.is_contiguous()
    B, P1 = scores.shape
    assert samples.shape == (B, P1, 3) and best.shape == (B, 3) and gt.shape == (B, 3) and out.shape == (B, 7)
    dr_min, dt_min = (float(remove[0]), float(remove[1])) if remove is not None else (0.0, 0.0)
    _lib.check(_lib.lib().snapb200_loc_nll(
        C.c_void_p(_ptr(scores)), C.c_void_p(_ptr(samples)), C.c_void_p(_ptr(best)), C.c_void_p(_ptr(gt)), B, P1,
        int(remove is not None), C.c_float(dr_min), C.c_float(dt_min), C.c_void_p(_ptr(out)),
        C.c_void_p(_ptr(dr_samples)), C.c_void_p(_ptr(dt_samples)), _stream()))


def sem_loss(logits: torch.Tensor, labels_area: torch.Tensor, valid_area: torch.Tensor,
             labels_excl: Optional[torch.Tensor], masks_indep: Optional[torch.Tensor], valid: torch.Tensor,
             num_area: int, num_excl: int, num_indep: int, weights: Optional[Sequence[Optional[torch.Tensor]]],
             out: torch.Tensor) -> None:
    """logits f32 [B, cells, ld]; labels i32 [B, cells]; masks u8; out f32 [B, SEM_OUT] (see include/snapb200.h)."""
    _require(logits, torch.float32, "logits")
    _require(labels_area, torch.int32, "labels_area")
    _require(valid_area, torch.uint8, "valid_area")
    _require(valid, torch.uint8, "valid")
    _require(out, torch.float32, "out")
    B, cells, ld = logits.shape
    assert logits.is_contiguous() and out.shape == (B, _lib.SEM_OUT)
    for t in (labels_area, valid_area, valid, labels_excl, masks_indep):
        assert t is None or t.is_contiguous()
    p = _lib.SemLossParams()
    p.B, p.cells, p.num_area, p.num_excl, p.num_indep, p.ld = B, cells, num_area, num_excl, num_indep, ld
    w = list(weights) if weights is not None else [None] * 4
    _lib.check(_lib.lib().snapb200_sem_loss(
        C.byref(p), C.c_void_p(_ptr(logits)), C.c_void_p(_ptr(labels_area)), C.c_void_p(_ptr(valid_area)),
        C.c_void_p(_ptr(labels_excl)), C.c_void_p(_ptr(masks_indep)), C.c_void_p(_ptr(valid)),
        C.c_void_p(_ptr(w[0])), C.c_void_p(_ptr(w[1])), C.c_void_p(_ptr(w[2])), C.c_void_p(_ptr(w[3])),
        C.c_void_p(_ptr(out)), _stream()))


# --------------------------------------------------------------------------------------------
# head-only training step (semantic fine-tuning, default 'mlp' decoder)
# --------------------------------------------------------------------------------------------
def sem_loss_grad(logits: torch.Tensor, labels_area: torch.Tensor, valid_area: torch.Tensor,
                  labels_excl: Optional[torch.Tensor], masks_indep: Optional[torch.Tensor], valid: torch.Tensor,
                  num_area: int, num_excl: int, num_indep: int, weights: Optional[Sequence[Optional[torch.Tensor]]],
                  counts: torch.Tensor, dlogits: torch.Tensor) -> None:
    """d mean_b(total_b) / d logits: logits f32 [B, cells, ld] -> dlogits bf16 [B*cells, ld_out]; counts f32 [B,2] scratch."""
    _require(logits, torch.float32, "logits")
    _require(dlogits, torch.bfloat16, "dlogits")
    _require(counts, torch.float32, "counts")
    B, cells, ld = logits.shape
    assert logits.is_contiguous() and dlogits.is_contiguous() and dlogits.shape[0] >= B * cells and counts.shape == (B, 2)
    p = _lib.SemLossParams()
    p.B, p.cells, p.num_area, p.num_excl, p.num_indep, p.ld = B, cells, num_area, num_excl, num_indep, ld
    w = list(weights) if weights is not None else [None] * 4
    _lib.check(_lib.lib().snapb200_sem_loss_grad(
        C.byref(p), C.c_void_p(_ptr(logits)), C.c_void_p(_ptr(labels_area)), C.c_void_p(_ptr(valid_area)),
        C.c_void_p(_ptr(labels_excl)), C.c_void_p(_ptr(masks_indep)), C.c_void_p(_ptr(valid)),
        C.c_void_p(_ptr(w[0])), C.c_void_p(_ptr(w[1])), C.c_void_p(_ptr(w[2])), C.c_void_p(_ptr(w[3])),
        C.c_void_p(_ptr(counts)), C.c_int(dlogits.shape[1]), C.c_void_p(_ptr(dlogits)), _stream()))


def relu_bwd(h: torch.Tensor, dx: torch.Tensor, elems: int) -> None:
    _require(h, torch.bfloat16, "h")
    _require(dx, torch.bfloat16, "dx")
    assert h.is_contiguous() and dx.is_contiguous() and h.numel() >= elems and dx.numel() >= elems
    _lib.check(_lib.lib().snapb200_relu_bwd(C.c_void_p(_ptr(h)), C.c_void_p(_ptr(dx)), C.c_longlong(elems), _stream()))


def dense_wgrad(x: torch.Tensor, dy: torch.Tensor, M: int, K: int, N: int, dW: torch.Tensor,
                db: Optional[torch.Tensor], workspace: Optional[torch.Tensor] = None) -> None:
    """dW f32 [K,N] = x[:M,:K]^T dy[:M,:N], db f32 [N] = column sums of dy (bf16 inputs, fp32 accumulation)."""
    _require(x, torch.bfloat16, "x")
    _require(dy, torch.bfloat16, "dy")
    _require(dW, torch.float32, "dW")
    assert x.stride(1) == 1 and dy.stride(1) == 1 and dW.is_contiguous() and dW.shape == (K, N)
    f = _lib.lib().snapb200_dense_wgrad_workspace
    f.restype = C.c_size_t
    need = int(f(C.c_longlong(M), K, N))
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    assert workspace.numel() * workspace.element_size() >= need
    _lib.check(_lib.lib().snapb200_dense_wgrad(
        C.c_void_p(_ptr(x)), C.c_longlong(x.stride(0)), C.c_void_p(_ptr(dy)), C.c_longlong(dy.stride(0)),
        C.c_longlong(M), K, N, C.c_void_p(_ptr(dW)), C.c_void_p(_ptr(db)), C.c_void_p(_ptr(workspace)),
        C.c_size_t(workspace.numel() * workspace.element_size()), _stream()))


def cast_pad_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    _require(src, torch.float32, "src")
    _require(dst, torch.bfloat16, "dst")
    rows, cols = src.shape
    assert src.is_contiguous() and dst.is_contiguous() and dst.shape[0] == rows and dst.shape[1] >= cols
    _lib.check(_lib.lib().snapb200_cast_pad_bf16(C.c_void_p(_ptr(src)), rows, cols, dst.shape[1], C.c_void_p(_ptr(dst)),
                                                 _stream()))


def adam_step(p: torch.Tensor, m: torch.Tensor, v: torch.Tensor, g: torch.Tensor, lr: float, step: int,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    for t, nm in ((p, "p"), (m, "m"), (v, "v"), (g, "g")):
        _require(t, torch.float32, nm)
        assert t.is_contiguous() and t.numel() == p.numel()
    _lib.check(_lib.lib().snapb200_adam_step(
        C.c_void_p(_ptr(p)), C.c_void_p(_ptr(m)), C.c_void_p(_ptr(v)), C.c_void_p(_ptr(g)), C.c_longlong(p.numel()),
        C.c_float(lr), C.c_float(b1), C.c_float(b2), C.c_float(eps), step, _stream()))


def sem_labels(sel_area: Sequence[Sequence[int]], sel_excl: Sequence[Sequence[int]], sel_indep: Sequence[int], num_gt: int,
               masks: torch.Tensor, bev_valid: torch.Tensor, labels_area: torch.Tensor, valid_area: torch.Tensor,
               labels_excl: Optional[torch.Tensor], masks_indep: Optional[torch.Tensor]) -> None:
    """masks u8 [rows, num_gt] -> labels / validity as consumed by `sem_loss` (label preparation on the device)."""
    _require(masks, torch.uint8, "masks")
    _require(bev_valid, torch.uint8, "bev_valid")
    rows = bev_valid.numel()
    assert masks.is_contiguous() and masks.numel() == rows * num_gt

    def table(sel):
        flat = []
        for s in sel:
            s = list(s)[:4]
            flat += s + [-1] * (4 - len(s))
        return (C.c_int * max(len(flat), 1))(*flat)
    ti = (C.c_int * max(len(sel_indep), 1))(*sel_indep)
    _lib.check(_lib.lib().snapb200_sem_labels(
        table(sel_area), len(sel_area), table(sel_excl), len(sel_excl), ti, len(sel_indep), num_gt,
        C.c_void_p(_ptr(masks)), C.c_void_p(_ptr(bev_valid)), C.c_longlong(rows), C.c_void_p(_ptr(labels_area)),
        C.c_void_p(_ptr(valid_area)), C.c_void_p(_ptr(labels_excl)), C.c_void_p(_ptr(masks_indep)), _stream()))


# --------------------------------------------------------------------------------------------
# backward building blocks of the 'resnet_stage' decoder (csrc/train_stage.cu)
# --------------------------------------------------------------------------------------------
def gn_backward(x: torch.Tensor, dy: torch.Tensor, n: int, H: int, W: int, Cc: int, acc: torch.Tensor,
                scale: torch.Tensor, bias: torch.Tensor, accb: torch.Tensor, dx: torch.Tensor, dscale: torch.Tensor,
                dbias: torch.Tensor, *, post_relu: bool = True, padded_out: bool = False,
                add: Optional[torch.Tensor] = None, pre_relu: bool = False, out_layout: Optional[int] = None,
                dy_phase: bool = False, dy_sub: Optional[torch.Tensor] = None) -> None:
    """GroupNorm(+ReLU) backward: x = the forward's GroupNorm input, acc = its statistics accumulators, dy = gradient
    w.r.t. the (activated) output; dx dense or zero-bordered; dscale / dbias f32 [Cc]; accb f64 scratch [n, Cc, 2]."""
    for t, nm in ((x, "x"), (dy, "dy"), (dx, "dx")):
        _require(t, torch.bfloat16, nm)
        assert t.is_contiguous()
    _require(acc, torch.float64, "acc")
    _require(accb, torch.float64, "accb")
    if out_layout is None:       # 0 dense, 1 zero-bordered (H+2, W+2), 2 bottom/right extended (H+1, W+1)
        out_layout = 1 if padded_out else 0
    opix = {0: H * W, 1: (H + 2) * (W + 2), 2: (H + 1) * (W + 1)}[out_layout]
    dpix = 4 * (H // 2 + 1) * (W // 2 + 1) if dy_phase else H * W
    assert accb.numel() >= n * Cc * 2 and x.numel() >= n * H * W * Cc and dy.numel() >= n * dpix * Cc
    assert dx.numel() >= n * opix * Cc
    if dy_sub is not None:
        _require(dy_sub, torch.bfloat16, "dy_sub")
        assert dy_sub.is_contiguous() and dy_sub.numel() >= n * (H // 2) * (W // 2) * Cc
    if add is not None:
        _require(add, torch.bfloat16, "add")
        assert add.is_contiguous() and add.numel() >= n * H * W * Cc
    for t, nm in ((scale, "scale"), (bias, "bias"), (dscale, "dscale"), (dbias, "dbias")):
        _require(t, torch.float32, nm)
        assert t.numel() >= Cc
    _lib.check(_lib.lib().snapb200_gn_backward(
        C.c_void_p(_ptr(x)), C.c_void_p(_ptr(dy)), C.c_void_p(_ptr(dy_sub)), int(dy_phase), C.c_void_p(_ptr(add)), n, H, W,
        Cc, C.c_void_p(_ptr(acc)), C.c_int(acc.stride(0)), C.c_void_p(_ptr(scale)), C.c_void_p(_ptr(bias)), int(pre_relu),
        int(post_relu), int(out_layout),
        C.c_void_p(_ptr(accb)), C.c_void_p(_ptr(dx)), C.c_void_p(_ptr(dscale)), C.c_void_p(_ptr(dbias)), _stream()))


def wt_segments(b_fwd: torch.Tensor, cout: int, cin: int, taps: int, out: torch.Tensor) -> None:
    """out bf16 [cin, taps*cout] <- forward GEMM operand b_fwd bf16 [>= cout, >= taps*cin] with every tap transposed."""
    _require(b_fwd, torch.bfloat16, "b_fwd")
    _require(out, torch.bfloat16, "out")
    assert b_fwd.stride(1) == 1 and out.stride(1) == 1 and b_fwd.shape[0] >= cout and out.shape[0] >= cin
    _lib.check(_lib.lib().snapb200_wt_segments(C.c_void_p(_ptr(b_fwd)), C.c_int(b_fwd.stride(0)), cout, cin, taps,
                                               C.c_void_p(_ptr(out)), C.c_int(out.stride(0)), _stream()))


def stdconv_backward(w: torch.Tensor, dws: torch.Tensor, dw: torch.Tensor) -> None:
    """w, dws, dw f32 [K, Cout]: gradient w.r.t. the raw kernel from the gradient w.r.t. the standardised kernel."""
    for t, nm in ((w, "w"), (dws, "dws"), (dw, "dw")):
        _require(t, torch.float32, nm)
        assert t.is_contiguous() and t.shape == w.shape and t.dim() == 2
    _lib.check(_lib.lib().snapb200_stdconv_backward(C.c_void_p(_ptr(w)), C.c_void_p(_ptr(dws)), w.shape[0], w.shape[1],
                                                    C.c_void_p(_ptr(dw)), _stream()))


# --------------------------------------------------------------------------------------------
# backward of the lift (csrc/lift_backward.cu)
# --------------------------------------------------------------------------------------------
def lift_gather_pool_backward(p: "_lib.LiftParams", views: torch.Tensor, fimg: torch.Tensor, xs: torch.Tensor,
                              ys: torch.Tensor, zs: torch.Tensor, dstats: torch.Tensor, gimg: torch.Tensor) -> None:
    """gimg f32 [V, Hf, Wf, D+S] += scatter-add of the cotangent of the statistics rows (zero gimg first)."""
    _require(fimg, torch.bfloat16, "fimg")
    _require(dstats, torch.bfloat16, "dstats")
    _require(gimg, torch.float32, "gimg")
    assert gimg.is_contiguous() and gimg.numel() == p.V * p.Hf * p.Wf * p.CF
    assert dstats.is_contiguous() and dstats.shape[1] == p.stats_ld and dstats.shape[0] >= p.X * p.Y * p.Z
    _lib.check(_lib.lib().snapb200_lift_gather_pool_backward(
        C.byref(p), C.c_void_p(_ptr(views)), C.c_void_p(_ptr(fimg)), C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)),
        C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(dstats)), C.c_void_p(_ptr(gimg)), _stream()))


def vertical_max_backward(vol: torch.Tensor, valid: torch.Tensor, dplane: torch.Tensor, cells: int, Z: int, Cc: int,
                          dvol: torch.Tensor) -> None:
    for t, nm in ((vol, "vol"), (dplane, "dplane"), (dvol, "dvol")):
        _require(t, torch.bfloat16, nm)
        assert t.is_contiguous()
    _require(valid, torch.uint8, "valid")
    assert vol.numel() >= cells * Z * Cc and dvol.numel() >= cells * Z * Cc and dplane.numel() >= cells * Cc
    _lib.check(_lib.lib().snapb200_vertical_max_backward(
        C.c_void_p(_ptr(vol)), C.c_void_p(_ptr(valid)), C.c_void_p(_ptr(dplane)), C.c_longlong(cells), Z, Cc,
        C.c_void_p(_ptr(dvol)), _stream()))


def match_head_backward(plane: torch.Tensor, valid: torch.Tensor, cells: int, Cc: int, kernel: torch.Tensor,
                        bias: torch.Tensor, dout: torch.Tensor, dy: torch.Tensor) -> None:
    """dy bf16 [cells, 32] = cotangent of the matching Dense output from dout = cotangent of mask(normalize(.))."""
    for t, nm in ((plane, "plane"), (dout, "dout"), (dy, "dy")):
        _require(t, torch.bfloat16, nm)
        assert t.is_contiguous()
    _require(kernel, torch.float32, "kernel")
    _require(bias, torch.float32, "bias")
    _require(valid, torch.uint8, "valid")
    assert kernel.shape == (Cc, 32) and kernel.is_contiguous() and dout.numel() >= cells * 32 and dy.numel() >= cells * 32
    _lib.check(_lib.lib().snapb200_match_head_backward(
        C.c_void_p(_ptr(plane)), C.c_void_p(_ptr(valid)), C.c_longlong(cells), Cc, C.c_void_p(_ptr(kernel)),
        C.c_void_p(_ptr(bias)), C.c_void_p(_ptr(dout)), C.c_void_p(_ptr(dy)), _stream()))


def fuse_max_backward(a: torch.Tensor, va: torch.Tensor, b: torch.Tensor, vb: Optional[torch.Tensor], dout: torch.Tensor,
                      cells: int, Cc: int, da: torch.Tensor, db: torch.Tensor) -> None:
    for t, nm in ((a, "a"), (b, "b"), (dout, "dout"), (da, "da"), (db, "db")):
        _require(t, torch.bfloat16, nm)
        assert t.is_contiguous() and t.numel() >= cells * Cc
    _lib.check(_lib.lib().snapb200_fuse_max_backward(
        C.c_void_p(_ptr(a)), C.c_void_p(_ptr(va)), C.c_void_p(_ptr(b)), C.c_void_p(_ptr(vb)), C.c_void_p(_ptr(dout)),
        C.c_longlong(cells), Cc, C.c_void_p(_ptr(da)), C.c_void_p(_ptr(db)), _stream()))


# --------------------------------------------------------------------------------------------
# backward of the sampling localizer's loss (csrc/localizer_backward.cu)
# --------------------------------------------------------------------------------------------
def loc_nll_backward(scores: torch.Tensor, remove: Optional[Sequence[float]], dr_samples: Optional[torch.Tensor],
                     dt_samples: Optional[torch.Tensor], dscores: torch.Tensor,
                     dtemperature: Optional[torch.Tensor] = None) -> None:
    """dscores f32 [B,P1] = d mean_b(nll_b) / d scores; remove = threshold_remove_accurate_poses (dr_min, dt_min) or None."""
    _require(scores, torch.float32, "scores")
    _require(dscores, torch.float32, "dscores")
    B, P1 = scores.shape
    assert scores.is_contiguous() and dscores.is_contiguous() and dscores.shape == (B, P1)
    dr_min, dt_min = (float(remove[0]), float(remove[1])) if remove is not None else (0.0, 0.0)
    _lib.check(_lib.lib().snapb200_loc_nll_backward(
        C.c_void_p(_ptr(scores)), C.c_void_p(_ptr(dr_samples)), C.c_void_p(_ptr(dt_samples)), B, P1,
        int(remove is not None), C.c_float(dr_min), C.c_float(dt_min), C.c_void_p(_ptr(dscores)),
        C.c_void_p(_ptr(dtemperature)), _stream()))


def loc_pose_scoring_backward(sim: torch.Tensor, point_scale: torch.Tensor, i_xy: torch.Tensor,
                              valid_j: Optional[torch.Tensor], poses: torch.Tensor, dscores: torch.Tensor, H: int, W: int,
                              cell_size: float, mask_out_of_bounds: bool, relu_mask: bool, dsim: torch.Tensor) -> None:
    """dsim bf16 [B, rows >= N, H*W] from dscores f32 [B,P] (the arguments of `loc_pose_scoring`)."""
    _require(sim, torch.bfloat16, "sim")
    _require(dsim, torch.bfloat16, "dsim")
    _require(poses, torch.float32, "poses")
    _require(dscores, torch.float32, "dscores")
    B, N = point_scale.shape
    P = poses.shape[1]
    assert poses.shape == (B, P, 3) and dscores.shape == (B, P) and poses.is_contiguous() and dscores.is_contiguous()
    assert sim.is_contiguous() and dsim.is_contiguous() and i_xy.is_contiguous()
    assert dsim.dim() == 3 and dsim.shape[0] == B and dsim.shape[1] >= N and dsim.shape[2] == H * W
    p = _lib.LocScoreParams()
    p.B, p.N, p.H, p.W, p.P = B, N, H, W, P
    p.cell_size = cell_size
    p.mask_out_of_bounds = int(mask_out_of_bounds)
    p.i_xy_batched = int(i_xy.dim() == 3)
    _lib.check(_lib.lib().snapb200_loc_pose_scoring_backward(
        C.byref(p), C.c_void_p(_ptr(sim)), C.c_void_p(_ptr(point_scale)), C.c_void_p(_ptr(i_xy)),
        C.c_void_p(_ptr(valid_j)), C.c_void_p(_ptr(poses)), C.c_void_p(_ptr(dscores)), int(relu_mask),
        int(dsim.shape[1]), C.c_void_p(_ptr(dsim)), _stream()))


def lift_select_pool_backward(p: "_lib.LiftParams", top_k: int, views: torch.Tensor, view_centers: torch.Tensor,
                              fimg: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor, zs: torch.Tensor,
                              dstats: torch.Tensor, gimg: torch.Tensor) -> None:
    """View-selection path (V > top_k): gimg f32 [V, Hf, Wf, D+S] += scatter-add of the statistics cotangent."""
    _require(fimg, torch.bfloat16, "fimg")
    _require(dstats, torch.bfloat16, "dstats")
    _require(gimg, torch.float32, "gimg")
    assert gimg.is_contiguous() and gimg.numel() == p.V * p.Hf * p.Wf * p.CF
    assert dstats.is_contiguous() and dstats.shape[1] == p.stats_ld and dstats.shape[0] >= p.X * p.Y * p.Z
    _lib.check(_lib.lib().snapb200_lift_select_pool_backward(
        C.byref(p), top_k, C.c_void_p(_ptr(views)), C.c_void_p(_ptr(view_centers)), C.c_void_p(_ptr(fimg)),
        C.c_void_p(_ptr(xs)), C.c_void_p(_ptr(ys)), C.c_void_p(_ptr(zs)), C.c_void_p(_ptr(dstats)),
        C.c_void_p(_ptr(gimg)), _stream()))


def upsample2x_backward(dy: torch.Tensor, n: int, h: int, w: int, Cc: int, dx: torch.Tensor) -> None:
    """dy bf16 [n, 2h, 2w, Cc] (cotangent of `upsample2x`'s output) -> dx bf16 [n, h, w, Cc]."""
    _require(dy, torch.bfloat16, "dy")
    _require(dx, torch.bfloat16, "dx")
    assert dy.is_contiguous() and dx.is_contiguous() and dy.numel() >= 4 * n * h * w * Cc and dx.numel() >= n * h * w * Cc
    _lib.check(_lib.lib().snapb200_upsample2x_backward(C.c_void_p(_ptr(dy)), n, h, w, Cc, C.c_void_p(_ptr(dx)), _stream()))


def maxpool3x3s2_backward(x: torch.Tensor, dy: torch.Tensor, n: int, H: int, W: int, Cc: int, dx: torch.Tensor) -> None:
    """x bf16 [n, H, W, Cc] (the pool's input), dy bf16 [n, (H-1)//2+1, (W-1)//2+1, Cc] -> dx bf16 [n, H, W, Cc]."""
    for t, nm in ((x, "x"), (dy, "dy"), (dx, "dx")):
        _require(t, torch.bfloat16, nm)
        assert t.is_contiguous()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    assert x.numel() >= n * H * W * Cc and dx.numel() >= n * H * W * Cc and dy.numel() >= n * Ho * Wo * Cc
    _lib.check(_lib.lib().snapb200_maxpool3x3s2_backward(C.c_void_p(_ptr(x)), C.c_void_p(_ptr(dy)), n, H, W, Cc,
                                                         C.c_void_p(_ptr(dx)), _stream()))
