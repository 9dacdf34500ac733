"""The torch gradient oracle of the lift backward (tests/lift_torch_ref.py) reproduces the NumPy oracle's statistics rows
(pinned against the reference's own StreetViewEncoder.__call__), and its autograd agrees with the closed forms the CUDA
kernel evaluates (tools/design/backward_formulas.py)."""
import os
import sys

import numpy as np
import torch

from lift_torch_ref import gather_pool_stats
from util import F, to_oracle_geometry

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "design"))
import backward_formulas as bf  # noqa: E402


def _scene(G=24, V=3, hw=(64, 96), seed=5, **layout):
    from oracle import bev_mapper as obm, grids as ogrids, streetview_encoder as osv
    from snap_b200 import params, synthetic
    rng = np.random.default_rng(seed)
    data = synthetic.make_tile(seed, V, hw, G, **layout)
    ocam, oT = to_oracle_geometry(data, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz, _ = obm.build_xyz_query(ogrids.Grid2D((G, G), 0.2), oT.t)
    fimg = rng.standard_normal((V, hw[0] // 4, hw[1] // 4, 160)).astype(F)
    fp = params.init_mlp(rng, 257, (256, 128))
    p2d, vis, depth, _ = osv.project_points_to_views(oT, ocam, xyz.reshape(-1, 3))
    return fimg, ocam, oT, xyz, fp, p2d, vis, depth


def test_torch_lift_forward_equals_numpy_oracle():
    from oracle import bev_mapper as obm
    fimg, ocam, oT, xyz, fp, p2d, vis, depth = _scene()
    dbg = {}
    obm.lift_scene(fimg, ocam, oT, xyz, fp, debug=dbg)
    ref = np.concatenate(dbg["stats"])
    got = gather_pool_stats(torch.from_numpy(fimg), p2d, vis, depth).numpy()
    assert got.shape == ref.shape and 0.02 < vis.any(-1).mean() < 0.95
    assert np.abs(got - ref).max() <= 2e-5 * (1 + np.abs(ref).max())


def test_torch_lift_autograd_equals_closed_forms():
    """One voxel seen by several views: autograd of the torch restatement == pool_multiview_backward +
    lift_gather_backward (the formulas of csrc/lift_backward.cu)."""
    fimg, ocam, oT, xyz, fp, p2d, vis, depth = _scene(V=4, seed=6, spacing=0.5, same_side=True)   # dense layout: overlapping views
    n = int(np.argmax(vis.sum(-1)))
    assert vis[n].sum() >= 2
    rng = np.random.default_rng(0)
    g = rng.standard_normal(257)
    t = torch.from_numpy(fimg.astype(np.float64)).requires_grad_(True)
    stats = gather_pool_stats(t, p2d[n:n + 1], vis[n:n + 1], depth[n:n + 1])
    (stats[0] * torch.from_numpy(g)).sum().backward()
    auto = t.grad.numpy()
    # closed forms
    V, Hf, Wf, CF = fimg.shape
    D, S = 128, 32
    pt = p2d[n].astype(F) - F(0.5)
    lo = np.floor(pt).astype(int)
    w1 = (pt - lo).astype(np.float64)
    feats, scores, taps_all, bins_all = np.zeros((V, D)), np.zeros(V), [], []
    for v in range(V):
        taps = [(min(max(lo[v, 0] + a, 0), Hf - 1), min(max(lo[v, 1] + b, 0), Wf - 1)) for a in range(2) for b in range(2)]
        wts = [(w1[v, 0] if a else 1 - w1[v, 0]) * (w1[v, 1] if b else 1 - w1[v, 1]) for a in range(2) for b in range(2)]
        f = sum(w * fimg[v, r, c].astype(np.float64) for (r, c), w in zip(taps, wts))
        c = np.log(np.clip(depth[n, v], 1, 32)) / np.log(32.0) * (S - 1)
        b0, b1, wb = int(np.clip(np.floor(c), 0, S - 1)), int(np.clip(np.floor(c) + 1, 0, S - 1)), c - np.floor(c)
        feats[v], scores[v] = f[:D], (1 - wb) * f[D + b0] + wb * f[D + b1]
        taps_all.append((taps, wts))
        bins_all.append((b0, b1, wb))
    df, ds = bf.pool_multiview_backward(feats, scores, vis[n], g[:D], g[D:2 * D], g[2 * D])
    closed = np.zeros_like(auto)
    for v in range(V):
        if vis[n, v]:
            taps, wts = taps_all[v]
            b0, b1, wb = bins_all[v]
            closed[v] = bf.lift_gather_backward((Hf, Wf, CF), taps, wts, (b0, b1), wb, df[v], ds[v], D)
    assert np.abs(closed - auto).max() <= 1e-6 * (1 + np.abs(auto).max())


def test_torch_select_lift_forward_close_to_numpy_oracle():
    """V > top_k: the torch restatement of the selective path (bf16 coordinate / weight arithmetic as constants, value
    roundings straight-through) against the NumPy oracle in bf16-emulation mode (which also rounds the values)."""
    from oracle import bev_mapper as obm, streetview_encoder as osv
    from util import bf16_np, rd_bf16, rel_l2
    from lift_torch_ref import gather_pool_stats_select
    fimg, ocam, oT, xyz, fp, p2d, vis, depth = _scene(V=6, seed=8, spacing=0.5, same_side=True)
    fimg = bf16_np(fimg)
    pts = xyz.reshape(-1, 3)
    idx, _ = osv.view_selection(pts, oT, vis, 4)
    g = lambda a: np.take_along_axis(a, idx[..., None] if a.ndim == 3 else idx, 1)
    dbg = {}
    obm.lift_scene(fimg, ocam, oT, xyz, fp, rd=rd_bf16, debug=dbg, top_k=4)
    ref = np.concatenate(dbg["stats"])
    assert np.array_equal(np.concatenate(dbg["view_indices"]), idx)
    got = gather_pool_stats_select(torch.from_numpy(fimg), g(p2d), idx, g(vis), g(depth)).numpy()
    seen = g(vis).any(-1)
    assert (g(vis).sum(-1) >= 2).mean() > 0.01 and not got[~seen].any() and not ref[~seen].any()
    assert rel_l2(got[seen, :128], ref[seen, :128]) < 1e-2 and rel_l2(got[seen, 128:256], ref[seen, 128:256]) < 3e-2
    assert np.abs(got[seen, 256] - ref[seen, 256]).max() < 5e-2
