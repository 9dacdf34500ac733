"""Backward kernels of the sampling localizer's loss (csrc/localizer_backward.cu) and the `LocalizerLossBackward` plan on
the GPU, against their torch emulation / torch autograd (tests/ops_emulation.py, tests/loc_torch_ref.py; both checked on
the CPU in tests/test_localizer_backward_plan_cpu.py).

First B200 run: round 2 (green at first run)."""
import numpy as np
import pytest
import torch

from util import F, bf16_np, rd_bf16, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("remove", [None, (1.5, 0.6)])
def test_loc_nll_backward_vs_emulation(remove):
    import ops_emulation as emu
    from snap_b200 import ops
    rng = np.random.default_rng(2)
    B, P1 = 3, 1001
    scores = torch.from_numpy((rng.standard_normal((B, P1)) * 3).astype(F))
    dr = torch.from_numpy(rng.uniform(0, 5, (B, P1)).astype(F))
    dt = torch.from_numpy(rng.uniform(0, 2, (B, P1)).astype(F))
    ds, dT = torch.zeros((B, P1), device="cuda"), torch.zeros(B, device="cuda")
    ops.loc_nll_backward(scores.cuda(), remove, dr.cuda() if remove else None, dt.cuda() if remove else None, ds, dT)
    rs, rT = torch.zeros((B, P1)), torch.zeros(B)
    emu.loc_nll_backward(scores, remove, dr, dt, rs, rT)
    assert np.abs(ds.cpu().numpy() - rs.numpy()).max() <= 1e-6 and np.abs(dT.cpu().numpy() - rT.numpy()).max() <= 1e-4
    assert abs(float(ds.sum())) < 1e-5                                    # soft-max minus one-hot sums to zero


@pytest.mark.parametrize("mask,relu", [(True, True), (False, False)])
def test_loc_pose_scoring_backward_vs_emulation(mask, relu):
    import ops_emulation as emu
    from snap_b200 import ops
    rng = np.random.default_rng(7)
    B, N, G, P, cell = 2, 50, 32, 300, 0.5
    sim = torch.from_numpy(bf16_np(rng.standard_normal((B, N, G * G)))).to(torch.bfloat16)
    if relu:
        sim = torch.relu(sim)
    ps = torch.from_numpy(np.where(rng.random((B, N)) < 0.8, 0.02, 0.0).astype(F))
    xy = torch.from_numpy(np.stack([rng.uniform(0.5, 6.0, N), rng.uniform(-3.0, 3.0, N)], -1).astype(F))
    poses = torch.from_numpy(np.stack([rng.uniform(-0.6, 0.6, (B, P)), rng.uniform(0.0, 8.0, (B, P)),
                                       rng.uniform(3.0, 13.0, (B, P))], -1).astype(F))
    vj = torch.from_numpy((rng.random((B, G * G)) < 0.9).astype(np.uint8))
    dscores = torch.from_numpy((rng.standard_normal((B, P)) * 0.1).astype(F))
    dsim = torch.full((B, N + 14, G * G), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.loc_pose_scoring_backward(sim.cuda(), ps.cuda(), xy.cuda(), vj.cuda() if mask else None, poses.cuda(), dscores.cuda(),
                                  G, G, cell, mask, relu, dsim)
    ref = torch.zeros((B, N + 14, G * G), dtype=torch.bfloat16)
    emu.loc_pose_scoring_backward(sim, ps, xy, vj if mask else None, poses, dscores, G, G, cell, mask, relu, ref)
    got, want = dsim[:, :N].float().cpu().numpy(), ref[:, :N].float().numpy()
    assert (dsim[:, N:].float() == 7.0).all(), "rows beyond N are not written"
    assert not got[ps.numpy() == 0].any() and np.abs(want).max() > 0
    assert rel_l2(got, want) < 1e-2


def test_localizer_loss_backward_chain_vs_autograd():
    """similarities and pose scores by the forward kernels, then `LocalizerLossBackward`; d f_q, d f_m and d temperature
    vs torch autograd of the same chain."""
    from loc_torch_ref import nll, pose_scores, pose_uv
    from snap_b200 import localizer_train, pose_estimation, types
    rng = np.random.default_rng(8)
    B, N, G, D, P1, cell, temp = 2, 48, 32, 32, 128, 0.5, 0.3
    remove = (1.5, 0.6)
    xy = np.stack([rng.uniform(0.5, 6.0, N), rng.uniform(-3.0, 3.0, N)], -1).astype(F)
    poses = np.stack([rng.uniform(-0.6, 0.6, (B, P1)), rng.uniform(0.0, 8.0, (B, P1)), rng.uniform(3.0, 13.0, (B, P1))], -1).astype(F)
    fq = bf16_np(rng.standard_normal((B, N, D)) / np.sqrt(D) * 2)
    fm = bf16_np(rng.standard_normal((B, G, G, D)) / np.sqrt(D) * 2)
    valid_pts = rng.random((B, N)) < 0.8
    valid_j = rng.random((B, G, G)) < 0.9
    dr, dt = rng.uniform(0, 5, (B, P1)).astype(F), rng.uniform(0, 2, (B, P1)).astype(F)
    cu = lambda a, dtp=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dtp) if dtp else torch.from_numpy(np.ascontiguousarray(a))).cuda()
    maps = pose_estimation.point_similarities(cu(fq, torch.bfloat16), cu(valid_pts.astype(np.uint8)), cu(fm, torch.bfloat16), temp, True)
    grid = types.Grid2D((G, G), cell)
    scores = pose_estimation.pose_scoring_many_batched(cu(poses), maps, cu(xy), cu(valid_j.astype(np.uint8)), grid, True)
    lb = localizer_train.LocalizerLossBackward(torch.device("cuda"))
    dfq, dfm, dtemp = lb.backward(maps, cu(fq, torch.bfloat16), cu(fm, torch.bfloat16), cu(xy), cu(valid_j.astype(np.uint8)),
                                  cu(poses), scores, cell, True, True, remove, cu(dr), cu(dt))
    torch.cuda.synchronize()
    T = torch.tensor(temp, requires_grad=True)
    total, qs, ms, ref_scores = 0, [], [], []
    for b in range(B):
        q, m = torch.from_numpy(fq[b]).requires_grad_(True), torch.from_numpy(fm[b].reshape(G * G, D)).requires_grad_(True)
        sim = torch.relu(rd_bf16(q @ m.T))
        sp = (sim * torch.exp(T) / max(int(valid_pts[b].sum()), 1)).reshape(N, G, G)
        sc = pose_scores(sp, pose_uv(poses[b], xy, cell), valid_pts[b], valid_j[b], True)
        removed = (dr[b] < remove[0]) & (dt[b] < remove[1])
        removed[0] = False
        total = total + nll(sc, removed) / B
        qs.append(q); ms.append(m); ref_scores.append(sc.detach().numpy())
    total.backward()
    assert np.abs(scores.cpu().numpy() - np.stack(ref_scores)).max() <= 2e-3 * (1 + np.abs(np.stack(ref_scores)).max())
    for b in range(B):
        eq, em = rel_l2(dfq[b].float().cpu().numpy(), qs[b].grad.numpy()), rel_l2(dfm[b].cpu().numpy(), ms[b].grad.numpy())
        print(f"example {b}: d f_q rel err {eq:.4f}, d f_m rel err {em:.4f}")
        assert eq < 3e-2 and em < 3e-2
    assert abs(float(dtemp.sum()) - float(T.grad)) <= 2e-2 * (1 + abs(float(T.grad)))
