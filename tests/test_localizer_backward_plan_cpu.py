"""Gradient oracle and launch plan of the localizer-loss backward on the CPU: tests/loc_torch_ref.py reproduces the NumPy
oracle's pose scores / NLL, and `snap_b200/localizer_train.py::LocalizerLossBackward` on the emulated operator layer
matches torch autograd of the whole chain features -> similarities -> pose scores -> NLL."""
import numpy as np
import torch

from loc_torch_ref import nll, pose_scores, pose_uv
from ops_emulation import emulated_ops
from util import F, bf16_np, rd_bf16


def _case(seed=3, N=40, G=16, P1=64, D=32):
    rng = np.random.default_rng(seed)
    cell = 0.5
    xy = np.stack([rng.uniform(0.5, 5.0, N), rng.uniform(-2.0, 2.0, N)], -1).astype(F)
    poses = np.stack([rng.uniform(-0.6, 0.6, P1), rng.uniform(0.0, 3.0, P1), rng.uniform(2.0, 6.0, P1)], -1).astype(F)
    fq = bf16_np(rng.standard_normal((N, D)) / np.sqrt(D) * 2)
    fm = bf16_np(rng.standard_normal((G * G, D)) / np.sqrt(D) * 2)
    valid_pts = rng.random(N) < 0.8
    valid_j = rng.random((G, G)) < 0.9
    return rng, cell, xy, poses, fq, fm, valid_pts, valid_j


def test_torch_scores_and_nll_equal_numpy_oracle():
    from oracle import grids as ogrids, pose_estimation as ope
    rng, cell, xy, poses, fq, fm, valid_pts, valid_j = _case()
    G = valid_j.shape[0]
    sim_points = np.maximum(fq @ fm.T, 0).reshape(len(fq), G, G).astype(F) * F(0.37)
    for mask in (False, True):
        ref = ope.pose_scoring_many(poses[:, 0], poses[:, 1:], sim_points, xy, valid_pts, valid_j, ogrids.Grid2D((G, G), cell), mask)
        got = pose_scores(torch.from_numpy(sim_points), pose_uv(poses, xy, cell), valid_pts, valid_j, mask).numpy()
        assert np.abs(got - ref).max() <= 1e-4 * (1 + np.abs(ref).max()) and np.abs(ref).max() > 0
    dr, dt = rng.uniform(0, 5, len(poses)).astype(F), rng.uniform(0, 2, len(poses)).astype(F)
    removed = (dr < 1.5) & (dt < 0.6)
    removed[0] = False
    sc = ref.astype(F)
    m = np.where(removed, -np.inf, sc)
    want = -(m[0] - m.max() - np.log(np.exp(m - m.max()).sum()))
    assert removed.any() and abs(float(nll(torch.from_numpy(sc), removed)) - want) < 1e-4


def test_localizer_loss_backward_plan_matches_autograd():
    from snap_b200 import localizer_train, pose_estimation
    B = 2
    cases = [_case(seed=5 + b) for b in range(B)]
    rng, cell = cases[0][0], cases[0][1]
    G, N, D, P1 = 16, 40, 32, 64
    scale, remove = float(np.exp(F(0.3))), (1.5, 0.6)
    dr = rng.uniform(0, 5, (B, P1)).astype(F)
    dt = rng.uniform(0, 2, (B, P1)).astype(F)
    # ---- reference: autograd of mean_b nll_b w.r.t. f_q, f_m and the temperature --------------------------------------
    T = torch.tensor(0.3, requires_grad=True)
    fqs, fms, total, scores_all, ps_all = [], [], 0, [], []
    for b, (_, _, xy, poses, fq, fm, valid_pts, valid_j) in enumerate(cases):
        q, m = torch.from_numpy(fq).requires_grad_(True), torch.from_numpy(fm).requires_grad_(True)
        sim = torch.relu(rd_bf16(q @ m.T))                                                  # bev_localizer.py:157-159
        w = 1.0 / max(int(valid_pts.sum()), 1)                                              # :170-172 (no confidences)
        sp = (sim * torch.exp(T) * w).reshape(N, G, G)
        sc = pose_scores(sp, pose_uv(poses, xy, cell), valid_pts, valid_j, True)
        removed = (dr[b] < remove[0]) & (dt[b] < remove[1])
        removed[0] = False
        total = total + nll(sc, removed) / B
        fqs.append(q); fms.append(m); scores_all.append(sc.detach()); ps_all.append(np.where(valid_pts, scale * w, 0).astype(F))
    total.backward()
    # ---- the product's plan on the emulated operator layer -------------------------------------------------------------
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    f_p_q = torch.stack([bf(c[4]) for c in cases])
    fmap = torch.stack([bf(c[5]).reshape(G, G, D) for c in cases])
    sim = torch.relu(torch.stack([(f_p_q[b].float() @ fmap[b].reshape(G * G, D).float().T).to(torch.bfloat16) for b in range(B)]))
    maps = pose_estimation.SimilarityMaps(sim=sim, scale=scale, point_scale=torch.from_numpy(np.stack(ps_all)),
                                          row_cdf=None, row_max=None, chunk_sum=None, row_sum=None, H=G, W=G)
    q_xy = torch.from_numpy(np.stack([c[2] for c in cases]))                                # batched points [B,N,2]
    vj = torch.from_numpy(np.stack([c[7] for c in cases]).astype(np.uint8))
    poses_t = torch.from_numpy(np.stack([c[3] for c in cases]))
    with emulated_ops():
        lb = localizer_train.LocalizerLossBackward(torch.device("cpu"))
        dfq, dfm, dtemp = lb.backward(maps, f_p_q, fmap, q_xy, vj, poses_t, torch.stack(scores_all), cell, True, True,
                                      remove, torch.from_numpy(dr), torch.from_numpy(dt))
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    for b in range(B):
        eq, em = rel(dfq[b].float().numpy(), fqs[b].grad.numpy()), rel(dfm[b].numpy(), fms[b].grad.numpy())
        print(f"example {b}: d f_q rel err {eq:.4f}, d f_m rel err {em:.4f}")
        assert eq < 2e-2 and em < 2e-2 and float(fqs[b].grad.norm()) > 1e-4
        assert not dfq[b].float().numpy()[~cases[b][6]].any(), "invalid query points receive no gradient"
    assert abs(float(dtemp.sum()) - float(T.grad)) <= 1e-3 * (1 + abs(float(T.grad)))
