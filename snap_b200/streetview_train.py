"""Backward of the camera -> BEV lift of one scene (all-views and view-selection paths, default statistics, 'max' vertical
pooling) from the
cotangent of the street-view feature plane down to the parameters of `fusion_mlp` / `proj_mlp` and to the encoder
features -- what `jax.grad` computes through `snap/models/streetview_encoder.py:228-286` and
`snap/models/bev_mapper.py:56-88`.

Launch plan (per scene; closed forms: tools/design/backward_formulas.py):

    plane  --vertical_max_backward-->  dvol [N,128]              (cotangent split evenly among tied maxima; 0 where invalid)
    fusion MLP (Dense 257->256, ReLU, Dense 256->128):           statistics and hidden rows are RECOMPUTED for the scene
        dW1 = hid^T dvol, dhid = dvol W1^T, ReLU mask            (the forward keeps one scene's rows at a time)
        dW0 = stats^T dhid, dstats = dhid W0^T                   split-K weight gradients, dX on the tcgen05 GEMM engine
    dstats --lift_gather_pool_backward-->  gimg f32 [V,Hf,Wf,160]   pooling + depth-score + bilinear-gather scatter-add
    proj MLP (ReLU, Dense 128->160): dWp = relu(crop)^T gimg, dcrop = (gimg Wp^T) * [crop > 0]

Parameters are the Flax kernels `[in, out]`; the gradients are fp32 arrays of the same layout, ACCUMULATED over the scenes
of a batch by the caller (`zero_grads`).  The image encoder below `dcrop` has no backward yet."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import image_encoder, ops

F = np.float32


class LiftBackward:
    def __init__(self, sv_params: Dict, device, stats_ld: int = 288, feature_dim: int = 128, num_scale_bins: int = 32):
        self.dev, self.ld, self.D, self.CF = device, stats_ld, feature_dim, feature_dim + num_scale_bins
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).copy()).to(device)
        k0 = np.zeros((stats_ld, 256), F)                       # [mean | var | score_max | zero padding] rows
        k0[: 2 * feature_dim + 1] = sv_params["fusion_mlp"]["Dense_0"]["kernel"]
        self.W = {"fusion_mlp/Dense_0/kernel": t(k0), "fusion_mlp/Dense_1/kernel": t(sv_params["fusion_mlp"]["Dense_1"]["kernel"]),
                  "proj_mlp/Dense_0/kernel": t(sv_params["proj_mlp"]["Dense_0"]["kernel"])}
        self.b = {"fusion_mlp/Dense_0/bias": t(sv_params["fusion_mlp"]["Dense_0"]["bias"]),
                  "fusion_mlp/Dense_1/bias": t(sv_params["fusion_mlp"]["Dense_1"]["bias"]),
                  "proj_mlp/Dense_0/bias": t(sv_params["proj_mlp"]["Dense_0"]["bias"])}
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        self.g = {k: z(*v.shape) for k, v in {**self.W, **self.b}.items()}      # accumulated over scenes
        self._g1 = {k: z(*v.shape) for k, v in {**self.W, **self.b}.items()}    # one scene
        # bf16 operands: forward B operands [out, in] and dX operands [in, out] (the Flax kernel as stored)
        self.Bf0 = self.W["fusion_mlp/Dense_0/kernel"].t().contiguous().to(torch.bfloat16)       # [256, ld]
        self.Bf1 = self.W["fusion_mlp/Dense_1/kernel"].t().contiguous().to(torch.bfloat16)       # [128, 256]
        self.Wc = {k: v.to(torch.bfloat16).contiguous() for k, v in self.W.items()}
        self._buf: Dict = {}

    def load_params(self, sv_params: Dict) -> None:
        """Refresh the device copies after an optimiser step (same shapes; buffers and gradient arrays are kept)."""
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(self.dev)
        self.W["fusion_mlp/Dense_0/kernel"][: 2 * self.D + 1].copy_(f(sv_params["fusion_mlp"]["Dense_0"]["kernel"]))
        self.W["fusion_mlp/Dense_1/kernel"].copy_(f(sv_params["fusion_mlp"]["Dense_1"]["kernel"]))
        self.W["proj_mlp/Dense_0/kernel"].copy_(f(sv_params["proj_mlp"]["Dense_0"]["kernel"]))
        for k in self.b:
            a, b, c = k.split("/")
            self.b[k].copy_(f(sv_params[a][b][c]))
        self.Bf0.copy_(self.W["fusion_mlp/Dense_0/kernel"].t())
        self.Bf1.copy_(self.W["fusion_mlp/Dense_1/kernel"].t())
        for k, v in self.W.items():
            self.Wc[k].copy_(v)

    def zero_grads(self) -> None:
        for v in self.g.values():
            v.zero_()

    def grads_tree(self) -> Dict:
        """Flax-named gradient tree (host): fusion_mlp / proj_mlp, kernels [in, out]."""
        out: Dict = {}
        for k, v in self.g.items():
            a, b, c = k.split("/")
            arr = v.cpu().numpy().copy()
            if k == "fusion_mlp/Dense_0/kernel":
                arr = arr[: 2 * self.D + 1]
            out.setdefault(a, {}).setdefault(b, {})[c] = arr
        return out

    def _buffers(self, N: int, rows_img: int, V: int, hf: int, wf: int) -> Dict:
        key = (N, rows_img)
        if key not in self._buf:
            bf = lambda r, c: torch.zeros((r, c), dtype=torch.bfloat16, device=self.dev)
            R = image_encoder._round_up(max(N, 128), 128)
            Ri = image_encoder._round_up(max(rows_img, 128), 128)
            self._buf[key] = dict(stats=bf(R, self.ld), hid=bf(R, 256), vol=bf(R, 128), dvol=bf(R, 128), dhid=bf(R, 256),
                                  dstats=bf(R, self.ld),
                                  valid=torch.zeros(R, dtype=torch.uint8, device=self.dev),
                                  gimg=torch.zeros((V, hf, wf, self.CF), dtype=torch.float32, device=self.dev),
                                  gimg_bf=bf(Ri, self.CF), dcrop=bf(Ri, self.D))
        return self._buf[key]

    def scene_backward(self, lp, views: torch.Tensor, fimg: torch.Tensor, crop: torch.Tensor, xs: torch.Tensor,
                       ys: torch.Tensor, zs: torch.Tensor, volume: Optional[torch.Tensor], valid: Optional[torch.Tensor],
                       dplane: torch.Tensor, top_k: Optional[int] = None, view_centers: Optional[torch.Tensor] = None,
                       max_view_distance: Optional[float] = None) -> torch.Tensor:
        """lp / views / fimg / xs / ys / zs: the arguments of the scene's forward `ops.lift_gather_pool`; crop bf16
        [V*hf*wf, 128] = relu(cropped finest FPN level) (the proj MLP's input); volume bf16 [N,128] / valid u8 [N] = the
        forward's feature volume, or None after the FUSED forward (`lift_fused_kernel` never materialises the volume): it is
        then recomputed here with one more GEMM, so training can keep the single-kernel forward; dplane bf16 [X*Y, 128].
        top_k / view_centers (/ max_view_distance): the scene ran on the view-selection path (V > top_k,
        `ops.lift_select_pool`), whose backward kernel repeats the selection.
        Adds this scene's parameter gradients to `self.g` and returns dcrop bf16 [V*hf*wf, 128] (cotangent of the
        un-activated encoder features)."""
        N, cells, Z = lp.X * lp.Y * lp.Z, lp.X * lp.Y, lp.Z
        rows_img = lp.V * lp.Hf * lp.Wf
        if N % 16 or rows_img % 16:
            raise NotImplementedError("voxel and texel counts must be multiples of 16 (split-K weight-gradient kernel)")
        buf, g = self._buffers(N, rows_img, lp.V, lp.Hf, lp.Wf), self._g1
        # recompute this scene's statistics and hidden rows (the forward keeps them for one scene at a time)
        select = top_k is not None and lp.V > top_k
        if select:
            ops.lift_select_pool(lp, top_k, max_view_distance, views, view_centers, fimg, xs, ys, zs, buf["stats"], buf["valid"])
        else:
            ops.lift_gather_pool(lp, views, fimg, xs, ys, zs, buf["stats"], buf["valid"])
        ops.gemm(buf["stats"], self.Bf0, buf["hid"], m_rows=N, seg_k=self.ld, bias=self.b["fusion_mlp/Dense_0/bias"], relu=True)
        if volume is None:      # fused forward: the volume rows of this scene again (Dense 256 -> 128, zero where invalid, :281-282)
            ops.gemm(buf["hid"], self.Bf1, buf["vol"], m_rows=N, bias=self.b["fusion_mlp/Dense_1/bias"], row_mask=buf["valid"])
            volume, valid = buf["vol"], buf["valid"]
        # vertical max -> volume rows (invalid voxels and non-maximal levels get 0)
        ops.vertical_max_backward(volume, valid, dplane, cells, Z, self.D, buf["dvol"])
        # fusion MLP
        ops.dense_wgrad(buf["hid"], buf["dvol"], N, 256, self.D, g["fusion_mlp/Dense_1/kernel"], g["fusion_mlp/Dense_1/bias"])
        ops.gemm(buf["dvol"], self.Wc["fusion_mlp/Dense_1/kernel"], buf["dhid"], m_rows=N, seg_k=self.D)
        ops.relu_bwd(buf["hid"], buf["dhid"], N * 256)
        ops.dense_wgrad(buf["stats"], buf["dhid"], N, self.ld, 256, g["fusion_mlp/Dense_0/kernel"], g["fusion_mlp/Dense_0/bias"])
        Wc0 = self.Wc["fusion_mlp/Dense_0/kernel"]             # [ld, 256]: N = 256 statistics columns + the rest, two launches
        ops.gemm(buf["dhid"], Wc0[:256], buf["dstats"][:, :256], m_rows=N, seg_k=256)
        ops.gemm(buf["dhid"], Wc0[256:], buf["dstats"][:, 256:], m_rows=N, seg_k=256)
        # pooling + depth score + bilinear gather: scatter-add into the projected feature maps
        buf["gimg"].zero_()
        if select:
            ops.lift_select_pool_backward(lp, top_k, views, view_centers, fimg, xs, ys, zs, buf["dstats"], buf["gimg"])
        else:
            ops.lift_gather_pool_backward(lp, views, fimg, xs, ys, zs, buf["dstats"], buf["gimg"])
        # proj MLP: fimg = relu(crop) Wp + bp
        ops.cast_pad_bf16(buf["gimg"].view(rows_img, self.CF), buf["gimg_bf"][:rows_img])
        ops.dense_wgrad(crop, buf["gimg_bf"], rows_img, self.D, self.CF, g["proj_mlp/Dense_0/kernel"], g["proj_mlp/Dense_0/bias"])
        ops.gemm(buf["gimg_bf"], self.Wc["proj_mlp/Dense_0/kernel"], buf["dcrop"], m_rows=rows_img, seg_k=self.CF)
        ops.relu_bwd(crop, buf["dcrop"], rows_img * self.D)
        for k in self.g:                                       # accumulation over the scenes of a batch (a few 100 KB)
            self.g[k] += g[k]
        return buf["dcrop"]


class MatchingHeadBackward:
    """Backward of `BEVMapper`'s matching head (`bev_mapper.py:284-291`: Dense 128 -> 32, L2 normalisation, mask) and of
    the modality fusion in front of it (`:225-252`, 'max' over street-view and aerial planes): from the cotangent of
    `bev_matching` to the gradients of `matching_proj` and the cotangents of the two modality planes (the street-view one
    feeds `LiftBackward.scene_backward`)."""

    def __init__(self, matching_proj: Dict, device):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).copy()).to(device)
        self.dev = device
        self.K, self.b = t(matching_proj["kernel"]), t(matching_proj["bias"])           # [C, 32], [32]
        self.C = self.K.shape[0]
        self.Kc = self.K.to(torch.bfloat16).contiguous()                                 # dX operand [N = C, K = 32]
        self.g = {"kernel": torch.zeros_like(self.K), "bias": torch.zeros_like(self.b)}
        self._g1 = {"kernel": torch.zeros_like(self.K), "bias": torch.zeros_like(self.b)}
        self._buf: Dict = {}

    def load_params(self, matching_proj: Dict) -> None:
        """Refresh the device copies after an optimiser step."""
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(self.dev)
        self.K.copy_(f(matching_proj["kernel"]))
        self.b.copy_(f(matching_proj["bias"]))
        self.Kc.copy_(self.K)

    def zero_grads(self) -> None:
        for v in self.g.values():
            v.zero_()

    def backward(self, plane: torch.Tensor, valid: torch.Tensor, dmatch: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
        """plane bf16 [cells, C] / valid u8 [cells] = `bev_features`; dmatch bf16 [cells, 32].  Writes (or, with
        `accumulate`, adds) the gradients of matching_proj into `self.g` and returns dplane bf16 [cells, C] (a buffer that
        the next call with the same number of cells overwrites)."""
        cells = valid.numel()
        if cells % 16:
            raise NotImplementedError("the number of BEV cells must be a multiple of 16 (split-K weight-gradient kernel)")
        if cells not in self._buf:
            R = image_encoder._round_up(max(cells, 128), 128)
            self._buf[cells] = (torch.zeros((R, 32), dtype=torch.bfloat16, device=self.dev),
                                torch.zeros((R, self.C), dtype=torch.bfloat16, device=self.dev))
        dy, dplane = self._buf[cells]
        ops.match_head_backward(plane, valid, cells, self.C, self.K, self.b, dmatch, dy)
        g = self._g1 if accumulate else self.g
        ops.dense_wgrad(plane, dy, cells, self.C, 32, g["kernel"], g["bias"])
        if accumulate:
            for k in self.g:
                self.g[k] += g[k]
        ops.gemm(dy, self.Kc, dplane, m_rows=cells, seg_k=32)
        return dplane

    def fusion_backward(self, sv_plane: torch.Tensor, sv_valid: torch.Tensor, aerial_plane: torch.Tensor,
                        dplane: torch.Tensor):
        """Cotangents (street-view, aerial) of the two modality planes from the cotangent of their valid-masked maximum
        (the aerial plane is valid everywhere, `bev_mapper.py:208-211`)."""
        cells = sv_valid.numel()
        da, db = torch.zeros_like(dplane), torch.zeros_like(dplane)
        ops.fuse_max_backward(sv_plane, sv_valid, aerial_plane, None, dplane, cells, self.C, da, db)
        return da, db
