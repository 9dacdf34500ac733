"""Differentiable torch restatement of the lift's float path for FIXED geometry (test infrastructure): bilinear gather
(streetview_encoder.py:69-76, grids.py:116-137), depth-score interpolation (:109-124) and weighted pooling (:141-178) of
one scene.  The projection (p2d, vis, depth) comes from the NumPy oracle and is treated as data.  Its forward is checked
against `oracle.bev_mapper.lift_scene`; its autograd is the gradient oracle of `snapb200_lift_gather_pool_backward`."""
import numpy as np
import torch

F = np.float32


def gather_pool_stats(fimg: torch.Tensor, p2d: np.ndarray, vis: np.ndarray, depth: np.ndarray, D: int = 128,
                      depth_min_max=(1.0, 32.0), rd=None) -> torch.Tensor:
    """fimg [V,Hf,Wf,D+S] (torch, may require grad); p2d [N,V,2] (row, col), vis [N,V], depth [N,V] -> stats [N, 2D+1].
    `rd` (optional): the half-precision materialisation of the interpolated maps and of the scores, at the points where
    `oracle.bev_mapper.lift_scene` applies it (gradient passes straight through the casts)."""
    if rd is None:
        rd = lambda t: t
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
    f = rd(f)
    feats, scales = f[..., :D], f[..., D:]
    mn, mx = depth_min_max
    d = torch.clamp(torch.from_numpy(depth.astype(F)), mn, mx)
    c = torch.log(d / mn) / float(np.log(F(mx / mn))) * (S - 1)
    blo = torch.floor(c)
    wb = c - blo
    b0 = torch.clamp(blo.long(), 0, S - 1)
    b1 = torch.clamp(blo.long() + 1, 0, S - 1)
    score = rd((1 - wb) * torch.gather(scales, -1, b0[..., None])[..., 0] + wb * torch.gather(scales, -1, b1[..., None])[..., 0])
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


def chain_forward(svp, enc, p2d, vis, depth, V, hf, wf, cells, Z, rd, tp=None, x=None, route_vol=None):
    """proj MLP -> gather / pooling -> fusion MLP -> mask (:282) -> vertical max (bev_mapper.py:80-86) as ONE autograd graph.
    Returns the leaves (parameters tp, encoder features x) and the forward tensors (crop, fimg, vol, plane, plane_valid)."""
    if tp is None:      # fresh leaves; pass `tp` to share the parameters between several scenes of one graph
        tp = {k: {n: {a: torch.from_numpy(np.ascontiguousarray(v, dtype=F)).requires_grad_(True) for a, v in d.items()}
                  for n, d in t.items()} for k, t in svp.items() if k in ("proj_mlp", "fusion_mlp")}
    if x is None:       # a leaf; pass `x` (e.g. the encoder's finest FPN level, [V*hf*wf, C]) to extend an existing graph
        x = torch.from_numpy(enc).requires_grad_(True)
    crop = torch.relu(x)
    fimg = rd(rd(crop @ tp["proj_mlp"]["Dense_0"]["kernel"]) + tp["proj_mlp"]["Dense_0"]["bias"])
    stats = rd(gather_pool_stats(fimg.reshape(V, hf, wf, -1), p2d, vis, depth, rd=rd))
    hid = torch.relu(rd(rd(stats @ tp["fusion_mlp"]["Dense_0"]["kernel"]) + tp["fusion_mlp"]["Dense_0"]["bias"]))
    vol = rd(rd(hid @ tp["fusion_mlp"]["Dense_1"]["kernel"]) + tp["fusion_mlp"]["Dense_1"]["bias"])
    valid = torch.from_numpy(vis.any(-1))
    vol = torch.where(valid[:, None], vol, torch.zeros(()))
    m = valid.reshape(cells, Z, 1)
    masked = torch.where(m, vol.reshape(cells, Z, -1), torch.full((), -float("inf")))
    plane = torch.where(m.any(1), masked.amax(1), torch.zeros(()))
    if route_vol is not None:
        # teacher-forced arg-max: the z level(s) that receive the cotangent are taken from ANOTHER forward's volume (the
        # CUDA one), ties split evenly like jnp.max / torch.amax; the forward value is unchanged up to near-ties.  Removes
        # the chaotic part of a GPU-vs-autograd comparison: a one-ulp difference between two near-tied levels otherwise
        # re-routes that cell's whole cotangent.
        rv = torch.from_numpy(np.ascontiguousarray(route_vol, dtype=F)).reshape(cells, Z, -1)
        rmask = torch.where(m, rv, torch.full((), -float("inf")))
        hit = (rmask == rmask.amax(1, keepdim=True)) & m
        wgt = hit.float() / hit.float().sum(1, keepdim=True).clamp(min=1.0)
        plane = (wgt * vol.reshape(cells, Z, -1)).sum(1)
    return tp, x, dict(crop=crop, fimg=fimg, vol=vol, plane=plane, plane_valid=m.any(1)[:, 0])


def chain_reference(svp, enc, p2d, vis, depth, V, hf, wf, cells, Z, dplane, rd, route_vol=None):
    """Autograd of loss = sum(plane * dplane) through `chain_forward`.  Returns the forward tensors the product's backward
    consumes (NumPy), the parameter gradients {proj_mlp, fusion_mlp} and the cotangent of the encoder features."""
    tp, x, t = chain_forward(svp, enc, p2d, vis, depth, V, hf, wf, cells, Z, rd, route_vol=route_vol)
    (t["plane"] * torch.from_numpy(dplane)).sum().backward()
    grads = {k: {n: {a: v.grad.numpy() for a, v in d.items()} for n, d in tt.items()} for k, tt in tp.items()}
    fwd = {k: t[k].detach().numpy() for k in ("crop", "fimg", "vol", "plane")}
    return fwd, grads, x.grad.numpy()


def _pool(feats, score, vis):
    """pool_multiview_features (:141-178), weighted branch: feats [N,K,D], score [N,K], vis [N,K] (numpy bool)."""
    v = torch.from_numpy(np.ascontiguousarray(vis))
    any_v = v.any(-1)
    v_ = torch.where(any_v[:, None], v, torch.ones_like(v))
    neg = torch.full_like(score, -float("inf"))
    mxs = torch.clamp(torch.where(v_, score, neg).amax(-1, keepdim=True), min=0.0)
    e = torch.where(v_, torch.exp(score - mxs), torch.zeros_like(score))
    w = e / e.sum(-1, keepdim=True)
    mean = (w[..., None] * feats).sum(1)
    var = (w[..., None] * (feats - mean[:, None]) ** 2).sum(1)
    smax = torch.where(v_, score, neg).amax(-1, keepdim=True)
    stats = torch.cat([mean, var, smax], -1)
    return torch.where(any_v[:, None], stats, torch.zeros_like(stats))


def gather_pool_stats_select(fimg: torch.Tensor, p2d: np.ndarray, idx: np.ndarray, vis: np.ndarray, depth: np.ndarray,
                             D: int = 128, depth_min_max=(1.0, 32.0)) -> torch.Tensor:
    """The V > top_k path: p2d [N,K,2], idx [N,K], vis [N,K], depth [N,K] = the GATHERED observations of the selected views
    (streetview_encoder.py:241-249); sampling = interpolate_views_selective (:80-105) with the coordinate / weight
    arithmetic in bf16 (constants here) and the value roundings straight-through."""
    rd = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16).float()
    V, Hf, Wf, CF = fimg.shape
    S = CF - D
    size = torch.tensor([Hf, Wf], dtype=torch.float32)
    pt = rd(p2d)
    pt = torch.clamp(torch.minimum((pt - 0.5).to(torch.bfloat16).float(), (size - 1).to(torch.bfloat16).float()), min=0.0)
    lo = torch.floor(pt)
    w_up = (pt - lo).to(torch.bfloat16).float()
    w_lo = (1.0 - w_up).to(torch.bfloat16).float()
    lo = lo.long()
    vi = torch.from_numpy(np.ascontiguousarray(idx)).long()
    f = 0
    for a in range(2):
        for b in range(2):
            r = torch.clamp(lo[..., 0] + a, 0, Hf - 1)
            c = torch.clamp(lo[..., 1] + b, 0, Wf - 1)
            w = ((w_up[..., 0] if a else w_lo[..., 0]) * (w_up[..., 1] if b else w_lo[..., 1])).to(torch.bfloat16).float()
            f = f + w[..., None] * fimg[vi, r, c]
    feats, scales = f[..., :D], f[..., D:]
    mn, mx = depth_min_max
    d = torch.clamp(torch.from_numpy(depth.astype(F)), mn, mx)
    c = torch.log(d / mn) / float(np.log(F(mx / mn))) * (S - 1)
    blo = torch.floor(c)
    wb = c - blo
    b0 = torch.clamp(blo.long(), 0, S - 1)
    b1 = torch.clamp(blo.long() + 1, 0, S - 1)
    score = (1 - wb) * torch.gather(scales, -1, b0[..., None])[..., 0] + wb * torch.gather(scales, -1, b1[..., None])[..., 0]
    return _pool(feats, score, vis)
