"""Restatement of snap/models/image_encoder.py. Test infrastructure (torch-CPU, NHWC fp32)."""
from __future__ import annotations

import math
from typing import Callable, Dict, List

import torch
import torch.nn.functional as Fnn

from . import resnet

_id = lambda a: a


def pad_to_multiple(images: torch.Tensor, stride: int) -> torch.Tensor:
    """image_encoder.py:32-39: pad bottom/right with zeros; pads a FULL stride when divisible (:37)."""
    h, w = images.shape[-3:-1]
    ph, pw = stride - h % stride, stride - w % stride
    return Fnn.pad(images, (0, 0, 0, pw, 0, ph))


def fpn_decoder(feats: List[torch.Tensor], p: Dict, rd: Callable = _id) -> List[torch.Tensor]:
    """image_encoder.py:53-94, norm='bit_resnet': relu -> GN -> 1x1 conv (no bias) -> + bilinear x2 of prev."""
    out, f_prev = [], None
    for level, f_skip in enumerate(feats):
        f = torch.relu(f_skip)
        f = resnet.group_norm(f, p[f"{level}_skip_norm"]["scale"], p[f"{level}_skip_norm"]["bias"], rd)
        f = resnet.conv(f, p[f"{level}_skip_conv"]["kernel"], rd=rd)
        if f_prev is not None:
            assert f.shape[1] == f_prev.shape[1] * 2 and f.shape[2] == f_prev.shape[2] * 2
            # jax.image.resize(..., 'bilinear') x2 == half-pixel bilinear with edge clamp
            up = Fnn.interpolate(f_prev.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False)
            f = rd(f + rd(up.permute(0, 2, 3, 1)))
        f_prev = f
        out.append(f)
    return out


def image_encoder(image: torch.Tensor, p: Dict, skip_root_block: bool = False, rd: Callable = _id, trace=None):
    """image_encoder.py:119-144.  image [B,H,W,3] in [0,1] -> (features coarse->fine (cropped), strides)."""
    h, w = image.shape[1:3]
    num_levels = sum(1 for k in p["encoder"] if k.startswith("block"))
    max_stride = (0 if skip_root_block else 2) + num_levels - 1  # :109-111
    padded = pad_to_multiple(rd(image), 2 ** max_stride)
    stages = resnet.resnet_v2(padded, p["encoder"], skip_root_block, rd, trace)
    skips = stages[::-1]  # :114, coarse -> fine
    outs = fpn_decoder(skips, p["decoder"], rd)
    feats, strides = [], []
    for f in outs:
        s = (padded.shape[1] / f.shape[1], padded.shape[2] / f.shape[2])
        hh, ww = int(round(math.ceil(h / s[0]))), int(round(math.ceil(w / s[1])))
        feats.append(f[:, :hh, :ww, :])
        strides.append(s)
    return feats, strides
