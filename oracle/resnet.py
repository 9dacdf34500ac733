"""Restatement of snap/models/resnet.py (BiT ResNet-v2) on torch-CPU fp32. Test infrastructure.

Tensors are torch NHWC float32.  `rd` is a rounding hook applied where the reference (run with a
half-precision `dtype`) would materialise an activation in that dtype; identity for fp32.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
import torch.nn.functional as Fnn

Tensor = torch.Tensor
_id = lambda a: a


def standardize(x: Tensor, dims, eps: float) -> Tensor:
    """resnet.py:34-41: x - mean; / sqrt(mean(x^2) + eps), in fp32."""
    x = x.float()
    x = x - x.mean(dim=dims, keepdim=True)
    return x / torch.sqrt((x * x).mean(dim=dims, keepdim=True) + eps)


def group_norm(x: Tensor, scale: Tensor, bias: Tensor, rd: Callable = _id, ngroups: int = 32) -> Tensor:
    """resnet.py:46-70.  x [N,H,W,C]; stats over (H,W,C/ngroups) per (image, group), eps=1e-5."""
    n, h, w, c = x.shape
    xg = x.reshape(n, h, w, ngroups, c // ngroups)
    xg = rd(standardize(xg, dims=[1, 2, 4], eps=1e-5))
    x = xg.reshape(n, h, w, c)
    x = rd(x * scale.reshape(1, 1, 1, c))
    return rd(x + bias.reshape(1, 1, 1, c))


def std_kernel(kernel: Tensor, rd: Callable = _id) -> Tensor:
    """StdConv, resnet.py:73-79: standardise the HWIO kernel over (kh,kw,in) per output channel, eps=1e-10."""
    return rd(standardize(kernel, dims=[0, 1, 2], eps=1e-10))


def conv(x: Tensor, kernel_hwio: Tensor, stride: int = 1, padding=0, rd: Callable = _id) -> Tensor:
    """flax nn.Conv on NHWC with an HWIO kernel, no bias.  padding: int pad or 'SAME' for 1x1."""
    w = kernel_hwio.permute(3, 2, 0, 1)
    if padding == "SAME":
        assert kernel_hwio.shape[0] == 1
        padding = 0  # 1x1 SAME: zero pad, out = ceil(in / stride)
    y = Fnn.conv2d(x.permute(0, 3, 1, 2), w, stride=stride, padding=padding)
    return rd(y.permute(0, 2, 3, 1))


def root_block(x: Tensor, p: Dict, rd: Callable = _id) -> Tensor:
    """resnet.py:82-100: 7x7/2 StdConv (pad 3) + 3x3/2 max-pool (pad 1, -inf)."""
    x = conv(x, std_kernel(p["conv_root"]["kernel"], rd), stride=2, padding=3, rd=rd)
    y = Fnn.max_pool2d(x.permute(0, 3, 1, 2), 3, stride=2, padding=1)
    return y.permute(0, 2, 3, 1)


def residual_unit(x: Tensor, p: Dict, stride: int, rd: Callable = _id) -> Tensor:
    """resnet.py:103-134 (pre-activation bottleneck)."""
    residual = x
    y = torch.relu(group_norm(x, p["gn1"]["scale"], p["gn1"]["bias"], rd))
    if "conv_proj" in p:  # resnet.py:121-122: projection consumes the PRE-ACTIVATED tensor
        residual = conv(y, std_kernel(p["conv_proj"]["kernel"], rd), stride=stride, padding="SAME", rd=rd)
    y = conv(y, std_kernel(p["conv1"]["kernel"], rd), rd=rd)
    y = torch.relu(group_norm(y, p["gn2"]["scale"], p["gn2"]["bias"], rd))
    y = conv(y, std_kernel(p["conv2"]["kernel"], rd), stride=stride, padding=1, rd=rd)
    y = torch.relu(group_norm(y, p["gn3"]["scale"], p["gn3"]["bias"], rd))
    y = conv(y, std_kernel(p["conv3"]["kernel"], rd), rd=rd)
    return rd(y + residual)


def resnet_stage(x: Tensor, p: Dict, first_stride: int, rd: Callable = _id, trace=None) -> Tensor:
    """resnet.py:137-155; returns the last unit's output.  `trace` (list) records (input, output) per unit."""
    names = sorted(k for k in p if k.startswith("unit"))
    for i, name in enumerate(names):
        y = residual_unit(x, p[name], first_stride if i == 0 else 1, rd)
        if trace is not None:
            trace.append((x, y))
        x = y
    return x


def resnet_v2(image: Tensor, p: Dict, skip_root_block: bool = False, rd: Callable = _id, trace=None):
    """resnet.py:184-216; returns [stage1, stage2, ...] outputs (last unit of each stage)."""
    x = rd(rd(image) * 2 - 1)  # resnet.py:199
    if skip_root_block:
        x = conv(x, std_kernel(p["conv_root"]["kernel"], rd), stride=1, padding=1, rd=rd)  # :200-208
    else:
        x = root_block(x, p["root_block"], rd)
    if trace is not None:
        trace.append((None, x))
    outs = []
    i = 1
    while f"block{i}" in p:
        x = resnet_stage(x, p[f"block{i}"], 1 if i == 1 else 2, rd, trace)
        outs.append(x)
        i += 1
    return outs
