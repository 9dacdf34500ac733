"""Exhaustive (x, y, theta) voting vs oracle/pose_exhaustive_voting.py."""
import numpy as np
import pytest
import torch

from util import F, assert_close_bf16, bf16_np, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=F))


def _planes(G, D, seed, B=1):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((B, G, G, D)); q /= np.linalg.norm(q, axis=-1, keepdims=True)
    m = rng.standard_normal((B, G, G, D)); m /= np.linalg.norm(m, axis=-1, keepdims=True)
    ii, jj = np.mgrid[:G, :G]
    ang = np.arctan2(jj - G / 2 + 0.5, ii - G * 0.1)
    qv = (np.abs(ang) < np.deg2rad(36)) & (ii > G * 0.1)            # 72 degree wedge
    mv = np.ones((G, G), bool); mv[: G // 8, : G // 5] = False        # a hole in the map
    qv = np.broadcast_to(qv, (B, G, G)).copy(); mv = np.broadcast_to(mv, (B, G, G)).copy()
    q = bf16_np(q * qv[..., None]); m = bf16_np(m * mv[..., None])
    return q, qv, m, mv


@pytest.mark.parametrize("G,R", [(32, 8), (64, 36)])
def test_templates_count_scores_vs_oracle(G, R):
    from oracle import grids as ogrids, pose_exhaustive_voting as opv
    from snap_b200 import ops, pose_exhaustive_voting as pv, types
    D, B = 32, 2
    q, qv, m, mv = _planes(G, D, 5, B)
    grid, ogrid = types.Grid2D((G, G), 0.2), ogrids.Grid2D((G, G), 0.2)
    dev = "cuda"
    qd, md = _t(q).to(torch.bfloat16).to(dev), _t(m).to(torch.bfloat16).to(dev)
    qvd, mvd = torch.from_numpy(qv.astype(np.uint8)).to(dev), torch.from_numpy(mv.astype(np.uint8)).to(dev)
    templates, t_valid = pv.sample_query_templates(qd, qvd, R, grid)
    scores = pv.template_matching(templates, t_valid, md, mvd)                 # map-row-major kernel ("rows")
    assert ops.xcorr_rows_supported(R, G)
    scores_v1 = pv.template_matching(templates, t_valid, md, mvd, kernel="gemm")  # segmented-GEMM kernel
    torch.cuda.synchronize()
    assert torch.equal(torch.isneginf(scores), torch.isneginf(scores_v1))
    fin_ = torch.isfinite(scores)
    assert (scores[fin_] - scores_v1[fin_]).abs().max() <= 1e-4 * scores_v1[fin_].abs().max()
    if ops.xcorr_sw_supported(R, G):
        scores_sw = pv.template_matching(templates, t_valid, md, mvd, kernel="sw")  # sliding window, N = 48
        torch.cuda.synchronize()
        assert torch.equal(torch.isneginf(scores), torch.isneginf(scores_sw))
        assert (scores[fin_] - scores_sw[fin_]).abs().max() <= 1e-4 * scores_v1[fin_].abs().max()
    for b in range(B):
        ot, otv = opv.sample_query_templates(q[b], qv[b], R, ogrid)
        assert np.array_equal(t_valid[b].cpu().numpy().astype(bool), otv), "template validity must be bit-exact"
        tref = pv.templates_to_reference_layout(templates, R)
        assert not templates[:, :, :, R:].any(), "padding rotations must be zero"
        assert_close_bf16(tref[b].float().cpu().numpy(), ot, "templates")
        # scores on IDENTICAL bf16 templates: only the fp32 summation order differs
        tb = tref[b].float().cpu().numpy()
        ref = opv.template_matching(tb, otv, m[b], mv[b])
        got = scores[b].cpu().numpy()
        assert np.array_equal(np.isneginf(got), np.isneginf(ref)), "-inf (min-overlap) mask must be bit-exact"
        fin = np.isfinite(ref)
        assert fin.any() and (~fin).any()
        scale = np.abs(ref[fin]).max()
        err = np.abs(got[fin] - ref[fin]).max()
        print(f"G={G} R={R} b={b}: max err {err:.3e} / scale {scale:.3e}, finite frac {fin.mean():.3f}")
        assert err <= 1e-3 * scale  # north_star: <= 1e-3 rel for correlation floats


def test_voting_end_to_end_and_peak():
    """exhaustive_pose_voting on a query cut from the map: the vote peaks at the known pose."""
    from oracle import grids as ogrids, pose_exhaustive_voting as opv
    from snap_b200 import pose_exhaustive_voting as pv, types
    G, R, D = 64, 36, 32
    q, qv, m, mv = _planes(G, D, 9, 1)
    q = m.copy(); qv = np.ones_like(qv)  # query == map, all valid -> identity pose
    mv[:] = True
    grid, ogrid = types.Grid2D((G, G), 0.2), ogrids.Grid2D((G, G), 0.2)
    dev = "cuda"
    pq = types.FeaturePlane(_t(q).to(torch.bfloat16).to(dev), torch.from_numpy(qv.astype(np.uint8)).to(dev))
    pm = types.FeaturePlane(_t(m).to(torch.bfloat16).to(dev), torch.from_numpy(mv.astype(np.uint8)).to(dev))
    scores = pv.exhaustive_pose_voting(pq, pm, R, grid)[0].cpu().numpy()
    ref = opv.exhaustive_pose_voting(q[0], qv[0], m[0], mv[0], R, ogrid)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(scores), fin)
    scale = np.abs(ref[fin]).max()
    print("e2e voting: max err / scale", np.abs(scores[fin] - ref[fin]).max() / scale)
    # templates are re-rounded to bf16 before the tensor-core contraction (the fp32 reference does not):
    assert np.abs(scores[fin] - ref[fin]).max() <= 2e-3 * scale
    k = np.unravel_index(np.argmax(np.where(fin, scores, -np.inf)), scores.shape)
    assert tuple(int(x) for x in k) == (0, G - 1, G - 1)
    idx = pv.exhaustive_tfm_to_index(*pv.exhaustive_index_to_tfm(np.array([3, 70, 41]), grid, R), grid, R)
    assert np.allclose(idx, [3, 70, 41], atol=1e-3)


def test_voting_full_size_properties():
    """G=128, R=36 (BASELINE config 4 shape): size-independent properties instead of an oracle run."""
    from snap_b200 import pose_exhaustive_voting as pv, types
    G, R, D = 128, 36, 32
    q, qv, m, mv = _planes(G, D, 3, 1)
    grid = types.Grid2D((G, G), 0.2)
    dev = "cuda"
    md, mvd = _t(m).to(torch.bfloat16).to(dev), torch.from_numpy(np.ones_like(mv).astype(np.uint8)).to(dev)
    templates, t_valid = pv.sample_query_templates(md, mvd, R, grid)
    # quadrant property (`pose_exhaustive_voting.py:63-68`): template k*R/4+r == rot90(template r, k, axes=(2,1))
    tq = pv.templates_to_reference_layout(templates, R)[0].float().cpu().numpy()
    for k in range(1, 4):
        assert np.array_equal(tq[k * (R // 4):(k + 1) * (R // 4)], np.rot90(tq[: R // 4], k, axes=(2, 1)))
    s1 = pv.template_matching(templates, t_valid, md, mvd)
    # linearity in the map: scores(2m) == 2 scores(m) exactly (power-of-two scaling is exact in bf16/fp32)
    s2 = pv.template_matching(templates, t_valid, md * 2, mvd)
    torch.cuda.synchronize()
    a, b = s1.cpu().numpy(), s2.cpu().numpy()
    fin = np.isfinite(a)
    assert np.array_equal(fin, np.isfinite(b)) and np.array_equal(2 * a[fin], b[fin])
    k = np.unravel_index(np.argmax(np.where(fin, a, -np.inf)), a.shape)
    assert tuple(int(x) for x in k[1:]) == (0, G - 1, G - 1)
    # all-valid map: no shift falls below the minimum overlap except where the template itself has too few valid cells
    assert fin.any()


def test_xcorr_rows_full_size_vs_direct_dot():
    """G=128, R=36, D=32 -- the shape of the 0.95-of-roofline headline (`xcorr_rows_kernel`) -- checked NUMERICALLY:
    12,000 randomly chosen outputs (r, u, v) against direct fp32 dot products of the SAME bf16 templates with the edge-padded
    map (`pose_exhaustive_voting.py:83-103`: S_r[u,v] = sum_ijd q_r[i,j,d] m_pad[u+i,v+j,d] / sum q_valid_r), NumPy on the
    CPU; and the -inf (minimum-overlap) mask of the WHOLE [36,255,255] volume bit-exact against an FFT convolution of the
    validity masks (integers <= 16384: exact after rounding).  Tolerance <= 1e-3 * max|ref| (north_star); measured 1e-6."""
    import scipy.signal
    from snap_b200 import pose_exhaustive_voting as pv, types
    G, R, D = 128, 36, 32
    q, qv, m, mv = _planes(G, D, 11, 1)
    grid = types.Grid2D((G, G), 0.2)
    dev = "cuda"
    qd, md = _t(q).to(torch.bfloat16).to(dev), _t(m).to(torch.bfloat16).to(dev)
    qvd, mvd = torch.from_numpy(qv.astype(np.uint8)).to(dev), torch.from_numpy(mv.astype(np.uint8)).to(dev)
    templates, t_valid = pv.sample_query_templates(qd, qvd, R, grid)
    scores = pv.template_matching(templates, t_valid, md, mvd, kernel="rows")
    torch.cuda.synchronize()
    got = scores[0].cpu().numpy()
    tq = pv.templates_to_reference_layout(templates, R)[0].float().cpu().numpy()      # [R,G,G,D], the kernel's bf16 operands
    tv = t_valid[0].cpu().numpy().astype(bool)
    U = 2 * G - 1
    # -inf mask of the whole volume: cnt = true convolution of the un-flipped template validity with the zero-padded map
    # validity, threshold 0.05 G^2 (`:100-101`)
    thr = F(0.05 * G * G)
    for r in range(R):
        cnt = np.rint(scipy.signal.fftconvolve(tv[r].astype(np.float64), mv[0].astype(np.float64), mode="full"))
        assert cnt.shape == (U, U)
        assert np.array_equal(np.isneginf(got[r]), cnt.astype(F) <= thr), f"-inf mask differs at rotation {r}"
    fin = np.isfinite(got)
    assert 0.2 < fin.mean() < 1.0
    m_pad = np.pad(m[0], ((G - 1, G - 1), (G - 1, G - 1), (0, 0)), mode="edge")       # `:83-85`
    den = tv.reshape(R, -1).sum(-1).astype(F)
    rng = np.random.default_rng(0)
    cand = np.argwhere(fin)
    pick = cand[rng.choice(len(cand), 12000, replace=False)]
    ref = np.empty(len(pick), F)
    for k, (r, u, v) in enumerate(pick):
        ref[k] = np.dot(tq[r].reshape(-1), m_pad[u:u + G, v:v + G].reshape(-1)) / den[r]
    val = got[pick[:, 0], pick[:, 1], pick[:, 2]]
    scale = float(np.abs(ref).max())
    err = float(np.abs(val - ref).max())
    print(f"xcorr_rows G=128 R=36: {len(pick)} sampled outputs, max |err| {err:.3e} / max |ref| {scale:.3e} = {err / scale:.2e}; "
          f"finite fraction {fin.mean():.3f}")
    from util import record_parity
    record_parity("xcorr_rows G=128 R=36", "12000 sampled (r,u,v) vs direct fp32 dot, max err / max |ref|", err / scale, 1e-3)
    assert err <= 1e-3 * scale


@pytest.mark.parametrize("G", [32, 64, 128])
def test_overlap_count_exact_generic_and_all_valid(G):
    """cnt = true convolution of the un-flipped template validity with the zero-padded map validity, exact
    integers; example 0 has a map with holes (popcount kernel), example 1 an all-valid map (rectangle-sum kernel)."""
    import scipy.signal
    from snap_b200 import ops
    rng = np.random.default_rng(G)
    B, R = 2, 4
    tv = rng.random((B, R, G, G)) < 0.4
    tv[0, 1] = False
    tv[1, 2] = True
    mv = rng.random((B, G, G)) < 0.8
    mv[1] = True
    U = 2 * G - 1
    cnt = torch.full((B, R, U, U), -1.0, dtype=torch.float32, device="cuda")
    den = torch.empty((B, R), dtype=torch.float32, device="cuda")
    ops.xcorr_count(torch.from_numpy(tv.astype(np.uint8)).cuda(), torch.from_numpy(mv.astype(np.uint8)).cuda(), cnt, den)
    got = cnt.cpu().numpy()
    assert np.array_equal(den.cpu().numpy(), tv.sum((-1, -2)).astype(F))
    for b in range(B):
        for r in range(R):
            ref = scipy.signal.convolve2d(tv[b, r].astype(np.int64), mv[b].astype(np.int64), mode="full")
            assert ref.shape == (U, U)
            assert np.array_equal(got[b, r], ref.astype(F)), (b, r, np.abs(got[b, r] - ref).max())
