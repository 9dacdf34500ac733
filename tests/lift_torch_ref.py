"""Differentiable torch restatement of the lift's float path for FIXED geometry (test infrastructure): bilinear gather
(streetview_encoder.py:69-76, grids.py:116-137), depth-score interpolation (:109-124) and weighted pooling (:141-178) of
one scene.  The projection (p2d, vis, depth) comes from the NumPy oracle and is treated as data.  Its forward is checked
against `oracle.bev_mapper.lift_scene`; its autograd is the gradient oracle of `snapb200_lift_gather_pool_backward`."""
import numpy as np
import torch

F = np.float32


def gather_pool_stats(fimg: torch.Tensor, p2d: np.ndarray, vis: np.ndarray, depth: np.ndarray, D: int = 128,
                      depth_min_max=(1.0, 32.0)) -> torch.Tensor:
    """fimg [V,Hf,Wf,D+S] (torch, may require grad); p2d [N,V,2] (row, col), vis [N,V], depth [N,V] -> stats [N, 2D+1]."""
    V, Hf, Wf, CF = fimg.shape
    S = CF - D
    N = p2d.shape[0]
    pt = torch.from_numpy((p2d.astype(F) - F(0.5)).astype(F))                       # grids.py:129
    lo = torch.floor(pt)
    w1 = pt - lo
    lo = lo.long()
    vi = torch.arange(V)[None, :].expand(N, V)
    f = 0
    for a in range(2):
        for b in range(2):
            r = torch.clamp(lo[..., 0] + a, 0, Hf - 1)
            c = torch.clamp(lo[..., 1] + b, 0, Wf - 1)
            w = (w1[..., 0] if a else 1 - w1[..., 0]) * (w1[..., 1] if b else 1 - w1[..., 1])
            f = f + w[..., None] * fimg[vi, r, c]                                     # [N, V, CF]
    feats, scales = f[..., :D], f[..., D:]
    mn, mx = depth_min_max
    d = torch.clamp(torch.from_numpy(depth.astype(F)), mn, mx)
    c = torch.log(d / mn) / float(np.log(F(mx / mn))) * (S - 1)
    blo = torch.floor(c)
    wb = c - blo
    b0 = torch.clamp(blo.long(), 0, S - 1)
    b1 = torch.clamp(blo.long() + 1, 0, S - 1)
    score = (1 - wb) * torch.gather(scales, -1, b0[..., None])[..., 0] + wb * torch.gather(scales, -1, b1[..., None])[..., 0]
    v = torch.from_numpy(np.ascontiguousarray(vis))
    any_v = v.any(-1)
    v_ = torch.where(any_v[:, None], v, torch.ones_like(v))                           # double-where (:150-152)
    neg = torch.full_like(score, -float("inf"))
    mxs = torch.clamp(torch.where(v_, score, neg).amax(-1, keepdim=True), min=0.0)    # softmax(where=, initial=0)
    e = torch.where(v_, torch.exp(score - mxs), torch.zeros_like(score))
    w = e / e.sum(-1, keepdim=True)
    mean = (w[..., None] * feats).sum(1)
    var = (w[..., None] * (feats - mean[:, None]) ** 2).sum(1)
    smax = torch.where(v_, score, neg).amax(-1, keepdim=True)
    stats = torch.cat([mean, var, smax], -1)
    return torch.where(any_v[:, None], stats, torch.zeros_like(stats))
