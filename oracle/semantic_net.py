"""Restatement of the semantic head forward of snap/models/semantic_net.py:145-198 (decoder_type='resnet_stage').
Test infrastructure."""
from __future__ import annotations

from typing import Callable, Dict

import numpy as np
import torch

from . import layers, resnet
from .bev_mapper import np_rd

F = np.float32
_id = lambda a: a


def semantic_decoder(features: np.ndarray, valid: np.ndarray, p: Dict, rd: Callable = _id) -> np.ndarray:
    """features [B,G,G,128], valid [B,G,G]; p = params['decoder'] with layers_0 (Dense), layers_1 (ResNetStage),
    layers_3 (MLP).  Returns logits f32 [B,G,G,K], zero where invalid (:185-186)."""
    rdn = np_rd(rd)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F))
    tt = lambda tree: {k: (tt(v) if isinstance(v, dict) else t(v)) for k, v in tree.items()}
    x = rdn(layers.dense(rdn(features), p["layers_0"]["kernel"], None))            # nn.Dense (:154-158)
    x = rdn(x + p["layers_0"]["bias"].astype(F))
    x = resnet.resnet_stage(t(x), tt(p["layers_1"]), 1, rd).numpy()                 # ResNetStage (:159), final output (:160)
    logits = layers.mlp(x, p["layers_3"], rd=rdn)                                    # MLP(dim, num_classes) (:161)
    return np.where(valid[..., None], logits.astype(F), F(0))
