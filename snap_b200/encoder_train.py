"""Backward of the FPN decoder of the image encoder (`snap/models/image_encoder.py:53-94`, norm='bit_resnet': per level
ReLU -> GroupNorm -> 1x1 conv (no bias) -> + x2 bilinear up-sampling of the coarser level): the first block of the encoder
backward.  Only the finest level is consumed downstream (`streetview_encoder.py:222`), so its cotangent
enters at the last level and walks towards the coarse ones through the up-sampling.

Per level, fine -> coarse (tensors as the forward plan `image_encoder.EncoderPlan.run_fpn` holds them):

    a   = GroupNorm(relu(x))                        recomputed (`snapb200_gn_apply`, pre_relu)
    dW  = a^T dout                                  `snapb200_dense_wgrad` over <= 1024-channel slices of a
    da  = dout W^T                                  GEMM engine, B = the transposed kernel (`snapb200_wt_segments`)
    dx  = GroupNorm backward, masked by x > 0       `snapb200_gn_backward(pre_relu=1, post_relu=0)`  (+ dscale, dbias)
    dout(coarser) = up-sampling backward of dout    `snapb200_upsample2x_backward`

The cotangents dx of the skip inputs (the stage outputs of the ResNet trunk) are returned: the trunk's own backward
(strided bottleneck units, root block) is not built yet."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import image_encoder, ops

F = np.float32


class FPNBackward:
    def __init__(self, decoder_params: Dict, n_img: int, shapes: List, device, out_dim: int = 128):
        """decoder_params: the Flax tree `image_encoder/decoder` ('{level}_skip_norm', '{level}_skip_conv'); shapes: per
        level (coarse -> fine) the (h, w, c) of the skip input."""
        self.n, self.dev, self.od = n_img, device, out_dim
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1).copy()).to(device)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        bank = self.bank = image_encoder._WeightBank(device)
        self.lv = []
        for level, (h, w, c) in enumerate(shapes):
            rows = n_img * h * w
            if rows % 16:
                raise NotImplementedError("n * h * w must be a multiple of 16 at every level (split-K weight-gradient kernel)")
            R = image_encoder._round_up(max(rows, 128), 128)
            self.lv.append(dict(h=h, w=w, c=c, rows=rows,
                                scale=f32(decoder_params[f"{level}_skip_norm"]["scale"]),
                                bias=f32(decoder_params[f"{level}_skip_norm"]["bias"]),
                                wk=bank.add(decoder_params[f"{level}_skip_conv"]["kernel"], False),
                                a=z(R, c, dt=torch.bfloat16), da=z(R, c, dt=torch.bfloat16), dx=z(R, c, dt=torch.bfloat16),
                                bt=z(c, out_dim, dt=torch.bfloat16), dup=z(R, out_dim, dt=torch.bfloat16),
                                accb=z(n_img, c, 2, dt=torch.float64),
                                g=dict(kernel=z(c, out_dim), scale=z(c), bias=z(c))))
        bank.finalize()

    def backward(self, skips: List[torch.Tensor], accs: List[torch.Tensor], dout_finest: torch.Tensor) -> List[torch.Tensor]:
        """skips[level] bf16 [n*h*w, c] (coarse -> fine) = the FPN inputs; accs[level] = the f64 GroupNorm accumulators of
        relu(skip) the forward used; dout_finest bf16 [n*h*w, out_dim] of the LAST level.  Fills the per-level gradients
        (`grads_tree`) and returns the cotangents of the skip inputs, coarse -> fine."""
        n, od = self.n, self.od
        self.bank.run()
        dout = dout_finest
        dskips = [None] * len(self.lv)
        for level in reversed(range(len(self.lv))):
            L = self.lv[level]
            h, w, c, rows = L["h"], L["w"], L["c"], L["rows"]
            ops.gn_apply(skips[level], n, h, w, c, accs[level], L["scale"], L["bias"], True, False, ops.LAYOUT_DENSE, L["a"])
            for c0 in range(0, c, 1024):                       # dW = a^T dout, <= 1024 input channels per launch
                kc = min(1024, c - c0)
                ops.dense_wgrad(L["a"][:, c0:c0 + kc], dout, rows, kc, od, L["g"]["kernel"][c0:c0 + kc], None)
            ops.wt_segments(self.bank.b_mats[L["wk"]], od, c, 1, L["bt"])
            ops.gemm(dout, L["bt"], L["da"], m_rows=rows, seg_k=od)
            ops.gn_backward(skips[level], L["da"], n, h, w, c, accs[level], L["scale"], L["bias"], L["accb"], L["dx"],
                            L["g"]["scale"], L["g"]["bias"], post_relu=False, pre_relu=True)
            dskips[level] = L["dx"]
            if level > 0:                                       # out = conv + upsample2x(coarser out): both get dout
                P = self.lv[level - 1]
                ops.upsample2x_backward(dout, n, P["h"], P["w"], od, P["dup"])
                dout = P["dup"]
        return dskips

    def grads_tree(self) -> Dict:
        out: Dict = {}
        for level, L in enumerate(self.lv):
            out[f"{level}_skip_norm"] = {"scale": L["g"]["scale"].cpu().numpy().reshape(1, 1, 1, -1).copy(),
                                         "bias": L["g"]["bias"].cpu().numpy().reshape(1, 1, 1, -1).copy()}
            out[f"{level}_skip_conv"] = {"kernel": L["g"]["kernel"].cpu().numpy().reshape(1, 1, L["c"], self.od).copy()}
        return out
