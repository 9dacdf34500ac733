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


# ------------------------------------------------------------------------------------------------------------------
# losses / metrics (snap/models/semantic_net.py:31-110, 254-343)
# ------------------------------------------------------------------------------------------------------------------
def balancing_weights(frequencies: Dict[str, float], classes, binary: bool = False, eps: float = 1e-3):
    """:31-53."""
    f = np.array([frequencies[c] for c in classes], dtype=np.float64)
    if not binary:
        f = f / f.sum()
    f = f.clip(min=eps)
    w = (1 / (f * len(classes))).astype(F)
    if binary:
        return w, (1 / ((1 - f).clip(min=eps) * len(classes))).astype(F)
    return w


def masked_mean(x: np.ndarray, mask: np.ndarray, axis):
    """layers.py:30-33."""
    div = np.sum(np.where(mask.any(axis, keepdims=True), mask, True), axis)
    return (np.sum(x * mask, axis, dtype=F) / div).astype(F)


def multiclass_crossentropy_metrics(logits: np.ndarray, labels: np.ndarray, valid: np.ndarray, weights=None):
    """:56-84.  optax.softmax_cross_entropy_with_integer_labels = logsumexp(logits) - logits[label].
    Returns (nll [B], accuracy [B], recall [B, N])."""
    l = logits.astype(F)
    m = l.max(-1, keepdims=True)
    lse = np.log(np.exp(l - m).sum(-1, dtype=F)).astype(F)
    nll = (lse - (np.take_along_axis(l, labels[..., None], -1)[..., 0] - m[..., 0])).astype(F)
    if weights is not None:
        nll = (nll * weights[labels]).astype(F)
    nll = masked_mean(nll, valid, (1, 2))
    mask = labels[..., None] == np.arange(l.shape[-1])
    correct = np.argmax(l, axis=-1) == labels
    acc = masked_mean(correct, valid, (1, 2))
    recall = masked_mean(correct[..., None], valid[..., None] & mask, (1, 2))
    return nll, acc, recall


def binary_crossentropy_metrics(logits: np.ndarray, gt_mask: np.ndarray, valid: np.ndarray, w_pos=None, w_neg=None):
    """:87-110.  optax.sigmoid_binary_cross_entropy = -y log_sigmoid(x) - (1 - y) log_sigmoid(-x).
    Returns (nll [B], recall [B, N])."""
    x = logits.astype(F)
    ls = lambda t: -(np.maximum(-t, 0) + np.log1p(np.exp(-np.abs(t)))).astype(F)
    y = gt_mask.astype(F)
    nll = (-y * ls(x) - (1 - y) * ls(-x)).astype(F)
    if w_pos is not None:
        nll = (nll * np.where(gt_mask, w_pos, w_neg)).astype(F)
    nll = masked_mean(nll.mean(-1, dtype=F), valid, (1, 2))
    correct = ((1 / (1 + np.exp(-x))) > 0.5) == gt_mask
    recall = masked_mean(correct, valid[..., None] & gt_mask, (1, 2))
    return nll, recall


def create_exclusive_labels(masks_all: np.ndarray, gt_classes, classes, add_void: bool = False):
    """:254-276: argmax over the selected ground-truth masks ('line' absorbs the other lane-marking classes)."""
    gi = {c: i for i, c in enumerate(gt_classes)}
    masks = masks_all[..., [gi[c] for c in classes]].copy()
    if "line" in classes:
        ml = masks_all[..., gi["line"]].copy()
        for c in ("stopline", "otherlanemarking"):
            if c in gi and c not in classes:
                ml |= masks_all[..., gi[c]]
        masks[..., list(classes).index("line")] = ml
    valid = masks.any(-1)
    labels = np.argmax(masks, -1)
    if add_void:
        labels = np.where(valid, labels, len(classes))
    return labels, valid


def loss_metrics(logits_areas, logits_excl, logits_indep, bev_valid, masks_all, gt_classes, area_classes, excl_classes,
                 indep_classes, area_frequencies=None, object_frequencies=None):
    """:300-343 (without transfer_labels_from_pcm).  Returns (losses dict of [B], metrics dict)."""
    la, va = create_exclusive_labels(masks_all, gt_classes, area_classes)
    wa = balancing_weights(area_frequencies, area_classes) if area_frequencies else None
    nll_a, acc_a, rec_a = multiclass_crossentropy_metrics(logits_areas, la, bev_valid & va, wa)
    losses = {"nll_areas": nll_a}
    metrics = {"accuracy": acc_a, "recall/average": rec_a.mean(-1), "recall_areas": rec_a}
    total = nll_a
    if logits_excl is not None:
        le, _ = create_exclusive_labels(masks_all, gt_classes, excl_classes, add_void=True)
        gi = {c: i for i, c in enumerate(gt_classes)}
        mi = masks_all[..., [gi[c] for c in indep_classes]]
        we = balancing_weights(object_frequencies, (*excl_classes, "void")) if object_frequencies else None
        nll_e, acc_e, rec_e = multiclass_crossentropy_metrics(logits_excl, le, bev_valid, we)
        wp, wn = balancing_weights(object_frequencies, indep_classes, binary=True) if object_frequencies else (None, None)
        nll_i, rec_i = binary_crossentropy_metrics(logits_indep, mi, bev_valid, wp, wn)
        total = ((total + (nll_e + nll_i) / 2) / 2).astype(F)
        losses.update(nll_objects_exclusive=nll_e, nll_objects_indep=nll_i)
        metrics.update({"accuracy/excl": acc_e, "recall/average/excl": rec_e.mean(-1), "recall_excl": rec_e,
                        "recall/average/indep": rec_i.mean(-1), "recall_indep": rec_i})
    losses["total"] = total
    return losses, metrics


# ------------------------------------------------------------------------------------------------------------------
# training loss of the head as a differentiable torch function (gradient oracle = torch autograd of this restatement)
# ------------------------------------------------------------------------------------------------------------------
def total_loss_torch(logits: "torch.Tensor", la, va, le, mi, bev_valid, num_area: int, num_excl: int,
                     w_area=None, w_excl=None, w_pos=None, w_neg=None):
    """mean over the batch (trainer.py:221) of `total` (:300-343) for logits f32 [B,H,W,K] (already masked, :186);
    labels / masks as NumPy arrays.  Same arithmetic as `loss_metrics` above (checked in tests/test_golden_semantics.py)."""
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    B = logits.shape[0]
    cells = logits.shape[1] * logits.shape[2]

    def mmean(x, mask):  # layers.masked_mean over (1, 2)
        m = t(mask)
        cnt = m.sum((1, 2))
        return (x * m).sum((1, 2)) / torch.where(cnt > 0, cnt, torch.full_like(cnt, float(cells)))

    def xent(l, labels, w):
        nll = torch.logsumexp(l, -1) - torch.gather(l, -1, t(labels, torch.int64)[..., None])[..., 0]
        return nll * t(w)[t(labels, torch.int64)] if w is not None else nll
    la_l, rest = logits[..., :num_area], logits[..., num_area:]
    total = mmean(xent(la_l, la, w_area), bev_valid & va)
    if rest.shape[-1] > 0:
        le_l, li_l = rest[..., :num_excl], rest[..., num_excl:]
        nll_e = mmean(xent(le_l, le, w_excl), bev_valid)
        y = t(mi)
        ls = torch.nn.functional.logsigmoid
        bce = -y * ls(li_l) - (1 - y) * ls(-li_l)
        if w_pos is not None:
            bce = bce * torch.where(y > 0, t(w_pos), t(w_neg))
        nll_i = mmean(bce.mean(-1), bev_valid)
        total = (total + (nll_e + nll_i) / 2) / 2
    return total.mean(), total


def mlp_head_forward_torch(features: "torch.Tensor", valid: np.ndarray, params: Dict, rd: Callable = _id):
    """semantic_net.py:147-152,185-186 in torch (differentiable): layers.MLP on the BEV features, f32 logits, zero where
    invalid.  params: {'Dense_i': {'kernel', 'bias'}} of torch tensors (requires_grad for the gradient oracle)."""
    x = features
    n = len(params)
    for i in range(n):
        if i > 0:
            x = torch.relu(x)
        x = rd(x @ params[f"Dense_{i}"]["kernel"])
        x = rd(x + params[f"Dense_{i}"]["bias"])
    v = torch.from_numpy(np.ascontiguousarray(valid))[..., None]
    return torch.where(v, x.float(), torch.zeros((), dtype=torch.float32))


def stage_head_forward_torch(features: "torch.Tensor", valid: np.ndarray, params: Dict, rd: Callable = _id):
    """`semantic_decoder` above (semantic_net.py:153-161,185-186: Dense -> ResNetStage -> MLP) in torch, differentiable
    w.r.t. `params` (torch tensors with requires_grad; dtype casts in `rd` pass the gradient straight through): the
    gradient oracle of the 'resnet_stage' head-only training step."""
    x = rd(features @ params["layers_0"]["kernel"])
    x = rd(x + params["layers_0"]["bias"])
    x = resnet.resnet_stage(x, params["layers_1"], 1, rd)
    x = mlp_head_forward_torch(x, valid, params["layers_3"], rd)
    return x
