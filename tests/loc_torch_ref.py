"""Differentiable torch restatement of the localizer's score / loss path for FIXED poses (test infrastructure):
similarities (bev_localizer.py:156-160), pose scoring (pose_estimation.py:50-85) and the NLL (bev_localizer.py:244-262).
Its forward is checked against the NumPy oracle; its autograd is the gradient oracle of csrc/localizer_backward.cu."""
import numpy as np
import torch

F = np.float32


def pose_uv(poses: np.ndarray, xy: np.ndarray, cell: float) -> np.ndarray:
    """poses [P,3] (angle, tx, ty), xy [N,2] -> uv [P,N,2] = (R p + t) / cell (geometry.Transform2D.transform)."""
    c, s = np.cos(poses[:, 0]).astype(F), np.sin(poses[:, 0]).astype(F)
    x = c[:, None] * xy[None, :, 0] - s[:, None] * xy[None, :, 1] + poses[:, 1:2]
    y = s[:, None] * xy[None, :, 0] + c[:, None] * xy[None, :, 1] + poses[:, 2:3]
    return (np.stack([x, y], -1) / F(cell)).astype(F)


def pose_scores(sim_points: torch.Tensor, uv: np.ndarray, valid_points: np.ndarray, valid_j, mask_out_of_bounds: bool):
    """sim_points [N,H,W] torch; uv [P,N,2]; -> scores [P] (interpolate_score_maps + the masked sum of pose_scoring)."""
    N, H, W = sim_points.shape
    p = torch.from_numpy(uv.astype(F))
    inb = ((p >= 0) & (p < torch.tensor([H, W], dtype=torch.float32))).all(-1)
    c = p - 0.5
    lo = torch.floor(c)
    wh = c - lo
    lo = lo.long()
    n_idx = torch.arange(N)[None, :].expand(p.shape[0], N)
    out, ok = 0, inb
    vj = None if valid_j is None else torch.from_numpy(np.ascontiguousarray(valid_j))
    for ci in (0, 1):
        for cj in (0, 1):
            r = torch.clamp(lo[..., 0] + ci, 0, H - 1)
            q = torch.clamp(lo[..., 1] + cj, 0, W - 1)
            w = (wh[..., 0] if ci else 1 - wh[..., 0]) * (wh[..., 1] if cj else 1 - wh[..., 1])
            out = out + w * sim_points[n_idx, r, q]
            if vj is not None:
                ok = ok & vj[r, q]
    vp = torch.from_numpy(np.ascontiguousarray(valid_points))[None, :].expand_as(ok)
    if mask_out_of_bounds:
        vp = vp & ok
    return torch.where(vp, out, torch.zeros(())).sum(-1)


def nll(scores: torch.Tensor, removed: np.ndarray) -> torch.Tensor:
    """bev_localizer.py:254-262 for one example: -log_softmax(scores with removed samples at -inf)[0]."""
    sc = torch.where(torch.from_numpy(removed), torch.full((), -float("inf")), scores)
    return torch.logsumexp(sc, 0) - sc[0]
