"""Sampling localizer (bev_localizer.py:156-218, pose_estimation.py) on the GPU against the oracle restatement
(oracle/pose_estimation.py, pinned by tests/test_golden_localizer.py), plus a full-size known-pose test."""
import numpy as np
import pytest
import torch

from util import F, bf16_np

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _unit(rng, *shape):
    x = rng.standard_normal(shape).astype(F)
    return bf16_np(x / np.linalg.norm(x, axis=-1, keepdims=True))


def _problem(seed, B=2, N=150, H=24, W=40, D=32, valid_frac=0.7):
    rng = np.random.default_rng(seed)
    fq = _unit(rng, B, N, D)
    fm = _unit(rng, B, H, W, D)
    # plant correspondences so that the soft-max is peaked somewhere
    for b in range(B):
        for n in range(0, N, 3):
            fm[b, rng.integers(H), rng.integers(W)] = fq[b, n]
    vq = rng.random((B, N)) < valid_frac
    fq = fq * vq[..., None]          # bev_matching features are zero where invalid (bev_mapper.py:289-291)
    vm = rng.random((B, H, W)) < 0.85
    i_xy = ((rng.random((N, 2)) - [0.5, 0.0]) * [6.0, 5.0]).astype(F)
    return rng, fq, fm, vq, vm, i_xy


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def _maps(fq, fm, vq, temperature, conf=None):
    from snap_b200 import pose_estimation as pe
    return pe.point_similarities(_dev(fq, torch.bfloat16), _dev(vq.astype(np.uint8)), _dev(fm, torch.bfloat16),
                                 temperature, True, _dev(conf) if conf is not None else None)


@pytest.mark.parametrize("with_conf", [False, True])
def test_similarities_softmax_and_point_weights(with_conf):
    from oracle import pose_estimation as ope
    rng, fq, fm, vq, vm, i_xy = _problem(1)
    B, N = vq.shape
    H, W = fm.shape[1:3]
    conf = rng.standard_normal((B, N)).astype(F) if with_conf else None
    if with_conf:
        vq[1] = False    # masked_softmax: an all-invalid mask becomes all-valid (layers.py:38-39)
        fq[1] = 0
    maps = _maps(fq, fm, vq, 2.0, conf)
    torch.cuda.synchronize()
    sim_g = maps.sim.float().cpu().numpy().reshape(B, N, H, W)
    for b in range(B):
        sim, prob = ope.point_similarities(fq[b], vq[b], fm[b], 2.0, True, conf[b] if with_conf else None, rd=bf16_np)
        # raw similarities: identical bf16 values up to fp32 accumulation-order flips of the rounding
        raw = bf16_np(np.maximum(np.einsum("nd,ijd->nij", fq[b], fm[b]), 0))
        assert (np.abs(sim_g[b] - raw) > 2 ** -7 * np.abs(raw) + 1e-6).mean() == 0
        assert (sim_g[b] != raw).mean() < 2e-3
        # sim_points = scale * w * sim, prob_points = w * softmax
        sp = maps.sim_points()[b].cpu().numpy()
        assert np.abs(sp - sim).max() <= 1e-3 * np.abs(sim).max()
        w = np.diff(np.concatenate([[0], maps.row_cdf[b].cpu().numpy()]))
        e = np.exp(sim_g[b].astype(np.float64) * maps.scale)
        Z = e.reshape(N, -1).sum(-1)
        Zg = (maps.row_sum[b].double() * torch.exp(maps.row_max[b].double())).cpu().numpy()
        assert np.abs(Zg / Z - 1).max() <= 1e-4
        cs = (maps.chunk_sum[b].double() * torch.exp(maps.row_max[b].double())[:, None]).cpu().numpy()
        assert np.abs(cs / e.sum(-1) - 1).max() <= 1e-4
        assert np.abs(w - prob.reshape(N, -1).sum(-1)).max() <= 2e-4 * prob.reshape(N, -1).sum(-1).max()
        ps = maps.point_scale[b].cpu().numpy()
        assert np.array_equal(ps != 0, vq[b] & (w > 0))


def test_sampling_matches_inverse_cdf_of_the_oracle():
    from oracle import pose_estimation as ope
    from snap_b200 import pose_estimation as pe
    rng, fq, fm, vq, vm, i_xy = _problem(2, B=2, N=90, H=20, W=36)
    B, N = vq.shape
    H, W = fm.shape[1:3]
    maps = _maps(fq, fm, vq, 2.0)
    K = 4000
    u = rng.random((B, K, 2)).astype(F)
    idx = pe.sample_correspondences(maps, _dev(u)).cpu().numpy()
    assert idx.min() >= 0 and (idx[..., 0] < N).all() and (idx[..., 1] < H).all() and (idx[..., 2] < W).all()
    sim_g = maps.sim.float().cpu().numpy().reshape(B, N, H, W)
    for b in range(B):
        # oracle: fp64 inverse CDF on the probabilities implied by the GPU's own bf16 similarities
        e = np.exp(sim_g[b].astype(np.float64) * maps.scale)
        prob = e / e.reshape(N, -1).sum(-1)[:, None, None] / max(int(vq[b].sum()), 1)
        ref = ope.sample_correspondences_inverse_cdf(prob, u[b])
        flat_g = (idx[b, :, 0].astype(np.int64) * H + idx[b, :, 1]) * W + idx[b, :, 2]
        flat_r = (ref[:, 0] * H + ref[:, 1]) * W + ref[:, 2]
        same = flat_g == flat_r
        # fp32 prefix sums vs fp64: a draw that lands within ~1e-6 of a cell boundary may pick the neighbour
        assert same.mean() >= 0.995, same.mean()
        # ... of the cell CDF (adjacent flat index) or of the point CDF (adjacent point)
        near = (np.abs(flat_g - flat_r) <= 1) | (np.abs(idx[b, :, 0] - ref[:, 0]) == 1)
        assert near[~same].all()
    # statistical check of the marginal over points: uniform over ALL points (prob_points is not masked, :170-172)
    cnt = np.bincount(idx[0, :, 0], minlength=N)
    assert cnt.min() > 0 and abs(cnt.mean() - K / N) < 1e-9 and cnt.max() < 4 * K / N


@pytest.mark.parametrize("retries", [1, 8])
def test_ransac_poses_against_oracle(retries):
    from oracle import grids, pose_estimation as ope
    from snap_b200 import pose_estimation as pe, types
    rng = np.random.default_rng(3)
    B, N, H, W, P = 2, 300, 128, 128, 500
    i_xy = ((rng.random((N, 2)) - [0.5, 0.0]) * [24.0, 16.0]).astype(F)
    idx = np.stack([rng.integers(0, N, (B, P * retries * 2)), rng.integers(0, H, (B, P * retries * 2)),
                    rng.integers(0, W, (B, P * retries * 2))], -1).astype(np.int32)
    idx[0, 6:8] = idx[0, 4:6]         # degenerate minimal sets: identical correspondences
    idx[1, 10, 1:] = idx[1, 11, 1:]   # same map cell, different points
    poses = pe.transforms_from_correspondences(_dev(idx), _dev(i_xy), P, retries, types.Grid2D((H, W), 0.2))
    poses = poses.cpu().numpy()
    g = grids.Grid2D((H, W), 0.2)
    for b in range(B):
        a, t = ope.sample_transforms_ransac(idx[b].astype(np.int64), i_xy, P, retries, g)
        d = np.abs((poses[b, :, 0] - a + np.pi) % (2 * np.pi) - np.pi)
        # the angle of a minimal set with baseline L is conditioned like eps/L: compare where it is well posed
        ok = d < 1e-4
        assert ok.mean() > 0.97, ok.mean()
        assert np.abs(poses[b, ok, 1:] - t[ok]).max() <= 5e-3
        assert np.isfinite(poses[b]).all()


@pytest.mark.parametrize("mask_oob", [False, True])
def test_pose_scoring_against_oracle(mask_oob):
    from oracle import grids, pose_estimation as ope
    from snap_b200 import pose_estimation as pe, types
    rng, fq, fm, vq, vm, i_xy = _problem(4, B=2, N=150, H=24, W=40)
    B, N = vq.shape
    H, W = fm.shape[1:3]
    maps = _maps(fq, fm, vq, 2.0)
    P = 700
    ang = rng.uniform(-np.pi, np.pi, (B, P)).astype(F)
    t = ((rng.random((B, P, 2)) * 1.6 - 0.3) * [H * 0.2, W * 0.2]).astype(F)
    ang[:, 0], t[:, 0] = 0, 0
    t[:, 1] = [1e6, -1e6]            # far away: clamped taps, no overflow
    poses = np.concatenate([ang[..., None], t], -1).astype(F)
    sc = pe.pose_scoring_many_batched(_dev(poses), maps, _dev(i_xy), _dev(vm.astype(np.uint8)), types.Grid2D((H, W), 0.2),
                                      mask_oob).cpu().numpy()
    sim_pts = maps.sim_points().cpu().numpy()
    for b in range(B):
        ref = ope.pose_scoring_many(ang[b], t[b], sim_pts[b], i_xy, vq[b], vm[b], grids.Grid2D((H, W), 0.2), mask_oob)
        tol = 1e-3 * np.abs(ref).max()
        bad = np.abs(sc[b] - ref) > tol
        # with the validity mask a point within 1 ulp of a cell border may flip (sincosf vs numpy); allow a handful
        assert bad.mean() <= (0.01 if mask_oob else 0.0), (bad.mean(), np.abs(sc[b] - ref).max(), tol)


def test_grid_refinement_against_oracle():
    from oracle import grids, pose_estimation as ope
    from snap_b200 import pose_estimation as pe, types
    rng, fq, fm, vq, vm, i_xy = _problem(5, B=2, N=40, H=48, W=48)
    B, N = vq.shape
    H, W = fm.shape[1:3]
    maps = _maps(fq, fm, vq, 2.0)
    init = np.array([[0.3, 4.0, 2.0], [-2.0, 5.0, 6.0]], F)
    refined, vol = pe.grid_refinement_batched(_dev(init), maps, _dev(i_xy), None, types.Grid2D((H, W), 0.2), False)
    refined, vol = refined.cpu().numpy(), vol.cpu().numpy()
    assert vol.shape == (B, 41, 41, 41)
    sim_pts = maps.sim_points().cpu().numpy()
    for b in range(B):
        a, t, ref = ope.grid_refinement(init[b, 0], init[b, 1:], sim_pts[b], i_xy, vq[b], vm[b], grids.Grid2D((H, W), 0.2), False)
        assert np.abs(vol[b] - ref).max() <= 1e-3 * np.abs(ref).max()
        k = np.unravel_index(np.argmax(vol[b]), vol[b].shape)
        assert ref[k] >= ref.max() - 1e-3 * np.abs(ref).max()
        if np.argmax(ref) == np.argmax(vol[b]):
            assert abs(refined[b, 0] - a) <= 1e-5 and np.abs(refined[b, 1:] - t).max() <= 1e-4


def test_argmax_and_loss_against_oracle():
    from oracle import pose_estimation as ope
    from snap_b200 import ops
    rng = np.random.default_rng(6)
    B, P1 = 3, 1001
    scores = (rng.standard_normal((B, P1)) * 3).astype(F)
    scores[1, 7] = scores[1, 400] = 50.0      # tie: first maximum wins
    scores[2, 0] = 60.0                       # the prepended ground truth is skipped by start = 1
    samples = np.concatenate([rng.uniform(-3.2, 3.2, (B, P1, 1)), rng.standard_normal((B, P1, 2)) * 3], -1).astype(F)
    gt = samples[:, 0].copy()
    samples[:, 1:9, 0] = gt[:, None, 0] + rng.uniform(-0.03, 0.03, (B, 8))
    samples[:, 1:9, 1:] = gt[:, None, 1:] + rng.uniform(-0.6, 0.6, (B, 8, 2))
    idx = torch.empty((B,), dtype=torch.int32, device="cuda")
    best = torch.empty((B, 3), dtype=torch.float32, device="cuda")
    ops.argmax_rows(_dev(scores), 1, idx, _dev(samples), best)
    ref_idx = np.argmax(scores[:, 1:], -1)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(best.cpu().numpy(), samples[np.arange(B), ref_idx + 1])
    for remove in (None, (1.5, 0.7)):
        out = torch.empty((B, 7), dtype=torch.float32, device="cuda")
        dr = torch.empty((B, P1), dtype=torch.float32, device="cuda")
        dt = torch.empty((B, P1), dtype=torch.float32, device="cuda")
        ops.loc_nll(_dev(scores), _dev(samples), best, _dev(gt), remove, out, dr, dt)
        out = out.cpu().numpy()
        for b in range(B):
            bp = best[b].cpu().numpy()
            nll, m, dr_s, dt_s = ope.loss_metrics(scores[b], samples[b, :, 0], samples[b, :, 1:], bp[0], bp[1:], gt[b, 0],
                                                  gt[b, 1:], remove)
            assert abs(out[b, 0] - nll) <= 1e-4 * (1 + abs(nll))
            assert abs(out[b, 1] - m["loc/err_max_rotation"]) <= 2e-3 and abs(out[b, 2] - m["loc/err_max_position"]) <= 1e-4
            assert bool(out[b, 3]) == m["loc/recall_top1"]
            assert np.abs(dr[b].cpu().numpy() - dr_s).max() <= 2e-3 and np.abs(dt[b].cpu().numpy() - dt_s).max() <= 1e-4
            for k, key in enumerate(["loc/recall_samples_0.5m_1", "loc/recall_samples_1m_2", "loc/recall_samples_2m_4"]):
                assert abs(out[b, 4 + k] - m[key]) <= 1.5 / (P1 - 1)


def test_query_points_lift_matches_grid_lift():
    """data['xy_bev'] (bev_mapper.py:163): lifting at arbitrary BEV points that coincide with cell centres of the
    regular grid gives exactly the planes of the regular grid at those cells (fused and unfused kernels)."""
    from snap_b200 import bev_mapper, configs, params, synthetic, types
    G, hw = 64, (224, 320)
    rng = np.random.default_rng(8)
    cfg = configs.bev_mapper(("streetview",))
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    data = synthetic.make_tile(51, 2, hw, G)
    grid = types.Grid2D((G, G), 0.2)
    cells = rng.permutation(G * G)[:700]
    ci, cj = cells // G, cells % G
    xy = np.stack([grid.cell_centers(0)[ci], grid.cell_centers(1)[cj]], -1)[:, None].astype(F)   # [N,1,2]
    for fused in (True, False):
        mapper = bev_mapper.BEVMapper(cfg, grid, fused_lift=fused)
        full = mapper.apply({"params": p}, dict(data))["bev_features"]
        ff, fv = full.features.float().cpu().numpy()[0], full.valid.cpu().numpy()[0]
        pts = mapper.apply({"params": p}, {**data, "xy_bev": xy})["bev_features"]
        pf, pv = pts.features.float().cpu().numpy()[0, :, 0], pts.valid.cpu().numpy()[0, :, 0]
        assert pf.shape == (700, 128)
        assert np.array_equal(pv, fv[ci, cj]) and 0.02 < pv.mean() < 0.98
        assert np.array_equal(pf, ff[ci, cj]), f"fused={fused}"


def test_per_example_xy_bev_matches_single_example_calls():
    """data['xy_bev'] [B,N,1,2] with DIFFERENT points per example (`bev_mapper.py:162-166`): BEVMapper.apply splits the
    batch into single-example launches; each example's result equals the result of calling it alone."""
    from snap_b200 import bev_mapper, configs, params, synthetic, types
    G, hw, B = 32, (96, 128), 2
    rng = np.random.default_rng(12)
    cfg = configs.bev_mapper(("streetview",))
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    data = synthetic.make_tile(52, 2, hw, G, batch=B)
    grid = types.Grid2D((G, G), 0.2)
    xy = np.stack([np.stack([rng.uniform(0.5, G * 0.2 - 0.5, 300), rng.uniform(0.5, G * 0.2 - 0.5, 300)], -1)[:, None]
                   for _ in range(B)]).astype(F)                                   # [B,300,1,2], different per example
    mapper = bev_mapper.BEVMapper(cfg, grid)
    both = mapper.apply({"params": p}, {**data, "xy_bev": xy})
    assert both["bev_features"].features.shape == (B, 300, 1, 128) and both["bev_matching"].features.shape == (B, 300, 1, 32)
    for b in range(B):
        cam, T = data["camera"], data["T_view2scene"]
        one = {"images": data["images"][b:b + 1], "camera": types.Camera(wh=cam.wh[b:b + 1], f=cam.f[b:b + 1], c=cam.c[b:b + 1]),
               "T_view2scene": types.Transform3D(R=T.R[b:b + 1], t=T.t[b:b + 1]), "xy_bev": xy[b]}
        alone = mapper.apply({"params": p}, one)
        for k in ("bev_features", "bev_matching"):
            assert torch.equal(alone[k].features[0], both[k].features[b]) and torch.equal(alone[k].valid[0], both[k].valid[b]), (k, b)
    assert 0.02 < float(both["bev_features"].valid.float().mean()) < 0.98


def test_matching_recovers_known_pose_full_size():
    """Config-4 sized matching block (4,652 frustum points, 128 x 128 map, D = 32, 10,000 poses x 8 retries, 41^3
    refinement) on discriminative synthetic planes: the query features are the map features at the cells the ground
    truth pose (90 degrees, integer-cell shift) maps the frustum points onto, so ~1 % of the minimal sets are exact
    and sampling -> retries -> Kabsch -> scoring -> argmax -> refinement must return the ground truth."""
    from snap_b200 import bev_localizer, configs, types
    G = 128
    rng = np.random.default_rng(10)
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov, cfg.num_pose_samples, cfg.num_pose_sampling_retries = True, 10_000, 8
    cfg.do_grid_refinement = True
    grid = types.Grid2D((G, G), 0.2)
    loc = bev_localizer.BEVLocalizer(cfg, None, grid)
    q = loc.q_xy_p[:, 0]
    N = len(q)
    qi = np.rint((q[:, 0] + 12.0) / 0.2 - 0.5).astype(int)
    qj = np.rint(q[:, 1] / 0.2 - 0.5).astype(int)
    gt = np.array([np.pi / 2, 20.0, 13.0], F)     # R(90) q + t: query point (i, j) -> map cell (99 - j, i + 5)
    fm = _unit(rng, 1, G, G, 32)
    vq = rng.random((1, N)) < 0.6
    fq = fm[0, 99 - qj, qi + 5][None] * vq[..., None]
    vm = np.ones((1, G, G), np.uint8)
    plane_q = types.FeaturePlane(_dev(fq.reshape(1, N, 1, 32), torch.bfloat16), _dev(vq.astype(np.uint8).reshape(1, N, 1)))
    plane_m = types.FeaturePlane(_dev(fm, torch.bfloat16), _dev(vm))
    Rz = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], F)
    T_q2m = types.Transform3D(R=Rz[None], t=np.array([[gt[1], gt[2], 0.0]], F))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(2)
    params = {"temperature": np.asarray(2.0, F)}
    pred = loc.match(params, plane_q, plane_m, None, T_q2m, {"sampling": gen})
    losses, metrics = loc.loss_metrics_function(pred, {"T_query2map": T_q2m}, params)
    torch.cuda.synchronize()
    sc = pred["scores_poses"][0].cpu().numpy()
    samples = pred["map_t_query_samples"][0].cpu().numpy()
    assert abs(samples[0, 0] - gt[0]) < 1e-6 and np.array_equal(samples[0, 1:], gt[1:])
    e2 = float(np.exp(F(2.0)))
    assert abs(sc[0] - e2) < 1e-3 * e2, "every valid point hits its own feature: score = exp(T) * mean(1)"
    d_ang = np.abs((samples[1:, 0] - gt[0] + np.pi) % (2 * np.pi) - np.pi)
    d_t = np.linalg.norm(samples[1:, 1:] - gt[1:], axis=-1)
    exact = (d_ang < 1e-4) & (d_t < 1e-2)
    print(f"matching: {int(exact.sum())} of 10000 sampled poses are the ground truth; gt score {sc[0]:.4f}, "
          f"best sample {sc[1:].max():.4f}, median sample {np.median(sc[1:]):.4f}, nll {losses['total'][0].item():.3f}")
    assert 20 <= exact.sum() <= 600, "expected ~1 % exact minimal sets (two exact correspondences in one of 8 retries)"
    ransac, refined = pred["map_t_query_ransac"][0].cpu().numpy(), pred["map_t_query"][0].cpu().numpy()
    assert exact[pred["best_index"][0].item()]
    assert abs(ransac[0] - gt[0]) < 1e-4 and np.abs(ransac[1:] - gt[1:]).max() < 1e-2
    assert abs(refined[0] - gt[0]) < 1e-4 and np.abs(refined[1:] - gt[1:]).max() < 1e-2, "zero offset wins the lattice"
    assert metrics["loc/err_max_position"][0].item() < 1e-2 and metrics["loc/err_max_rotation"][0].item() < 1e-2
    rec = metrics["loc/recall_samples_0.5m_1°"][0].item()
    assert exact.mean() - 1e-6 <= rec <= exact.mean() + 0.05
    vol = pred["scores_grid_refine"][0].cpu().numpy()
    assert vol.shape == (41, 41, 41) and np.unravel_index(np.argmax(vol), vol.shape) == (20, 20, 20)


def test_localizer_end_to_end_full_size():
    """BASELINE configs[3]-sized localization through BEVLocalizer.apply with random-init encoders: map tile (4 views,
    G = 128), query = one of its views in its own frame (4,652 field-of-view points through data['xy_bev']), 10,000
    poses x 8 retries + 41^3 refinement.  Random-init features are nearly constant over the scene (every similarity is
    ~0.96), so no pose is recoverable here; the test checks the plumbing and, at full size, the scores of a subset of
    poses against the oracle evaluated on the GPU's own planes."""
    from oracle import grids as ogrids, pose_estimation as ope
    from snap_b200 import bev_localizer, configs, params, synthetic, types
    G, hw = 128, (480, 640)
    rng = np.random.default_rng(9)
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov = True
    cfg.num_pose_samples = 10_000
    cfg.num_pose_sampling_retries = 8
    cfg.do_grid_refinement = True
    grid = types.Grid2D((G, G), 0.2)
    loc = bev_localizer.BEVLocalizer(cfg, None, grid)
    mp = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg.bev_mapper)))
    variables = {"params": loc.init_params(mp)}
    data = synthetic.make_tile(61, 4, hw, G)
    v = 2                                             # an even view: looks along +y of the map frame
    T = data["T_view2scene"]
    cam_xy = T.t[0, v, :2]
    t_q2m = (np.round(cam_xy / 0.2) * 0.2).astype(F)  # query origin on a cell corner: frustum points = map cell centres
    z_off = (np.median(T.t[..., -1].astype(F), axis=-1).astype(F) - F(4.0)).astype(F)
    qT = types.Transform3D(R=T.R[:, [v]].copy(), t=(T.t[:, [v]] - np.array([t_q2m[0], t_q2m[1], 0], F)).astype(F))
    cam = data["camera"]
    query = {"images": np.ascontiguousarray(data["images"][:, [v]]),
             "camera": types.Camera(wh=cam.wh[:, [v]].copy(), f=cam.f[:, [v]].copy(), c=cam.c[:, [v]].copy()),
             "T_view2scene": qT, "z_offset": z_off}
    T_q2m = types.Transform3D(R=np.eye(3, dtype=F)[None], t=np.array([[t_q2m[0], t_q2m[1], 0.0]], F))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    pred = loc.apply(variables, {"map": {**data, "z_offset": z_off}, "query": query, "T_query2map": T_q2m},
                     rngs={"sampling": gen})
    losses, metrics = loc.loss_metrics_function(pred, {"T_query2map": T_q2m}, variables["params"])
    torch.cuda.synchronize()
    vq = pred["query"]["bev_matching"].valid
    assert vq.shape == (1, 4652, 1) and vq.sum().item() > 300
    sc = pred["scores_poses"][0].cpu().numpy()
    assert sc.shape == (10_001,) and np.isfinite(sc).all()
    best = pred["map_t_query"][0].cpu().numpy()
    ransac = pred["map_t_query_ransac"][0].cpu().numpy()
    idx = pred["correspondences"][0].cpu().numpy()
    assert idx.shape == (160_000, 3) and idx.min() >= 0 and (idx[:, 0] < 4652).all() and (idx[:, 1:] < G).all()
    print(f"localizer: valid query points {int(vq.sum())}, gt score {sc[0]:.4f}, best sample {sc[1:].max():.4f}, "
          f"ransac pose {ransac}, refined {best}, gt {t_q2m}, nll {losses['total'][0].item():.3f}")
    assert pred["best_index"][0].item() == int(np.argmax(sc[1:]))
    assert np.array_equal(ransac, pred["map_t_query_samples"][0, 1 + int(np.argmax(sc[1:]))].cpu().numpy())
    assert abs(losses["total"][0].item() + (sc[0] - sc.max() - np.log(np.exp(sc - sc.max()).sum()))) < 1e-3
    err_t = float(np.linalg.norm(best[1:] - t_q2m))
    assert abs(metrics["loc/err_max_position"][0].item() - err_t) < 1e-3
    vol = pred["scores_grid_refine"][0].cpu().numpy()
    assert vol.shape == (41, 41, 41) and np.isfinite(vol).all()
    assert vol.max() >= sc[1:].max() - 1e-4 * abs(sc[0]), "the lattice contains the initial pose (zero offset)"
    assert np.isfinite(best).all()
    # full-size parity of pose_scoring on a subset of the sampled poses, oracle on the GPU's own planes
    maps = pred["similarity_maps"]
    sim_pts = maps.sim_points()[0].cpu().numpy()
    sub = np.r_[0:24, 5000:5024]
    samples = pred["map_t_query_samples"][0].cpu().numpy()
    ref = ope.pose_scoring_many(samples[sub, 0], samples[sub, 1:], sim_pts, loc.q_xy_p[:, 0], vq[0, :, 0].cpu().numpy().astype(bool),
                                pred["map"]["bev_matching"].valid[0].cpu().numpy().astype(bool), ogrids.Grid2D((G, G), 0.2), False)
    assert np.abs(sc[sub] - ref).max() <= 1e-3 * np.abs(ref).max(), np.abs(sc[sub] - ref).max()


def test_pose_scoring_edge_cases():
    """a 256 x 256 map (one similarity map no longer fits twice in shared memory: single-buffer path), no valid point
    at all (scores are exactly zero), a pose count that is not a multiple of the chunk size, batched query points."""
    from oracle import grids, pose_estimation as ope
    from snap_b200 import pose_estimation as pe, types
    rng, fq, fm, vq, vm, i_xy = _problem(7, B=1, N=24, H=256, W=256)
    maps = _maps(fq, fm, vq, 2.0)
    P = 2049 + 17
    ang = rng.uniform(-np.pi, np.pi, (1, P)).astype(F)
    t = (rng.random((1, P, 2)) * 51.2).astype(F)
    poses = np.concatenate([ang[..., None], t], -1).astype(F)
    i_xy_b = np.ascontiguousarray(i_xy[None])                      # [B,N,2]: per-example query points
    g256 = types.Grid2D((256, 256), 0.2)
    sc = pe.pose_scoring_many_batched(_dev(poses), maps, _dev(i_xy_b), None, g256, False).cpu().numpy()
    ref = ope.pose_scoring_many(ang[0], t[0], maps.sim_points()[0].cpu().numpy(), i_xy, vq[0], vm[0],
                                grids.Grid2D((256, 256), 0.2), False)
    assert np.abs(sc[0] - ref).max() <= 1e-3 * np.abs(ref).max()
    # no valid point: num_valid clips to 1, every point is masked -> all scores 0 (pose_estimation.py:78-84)
    vq0 = np.zeros_like(vq)
    maps0 = _maps(fq * 0, fm, vq0, 2.0)
    sc0 = pe.pose_scoring_many_batched(_dev(poses), maps0, _dev(i_xy), None, g256, False).cpu().numpy()
    assert np.array_equal(sc0, np.zeros_like(sc0))
    assert np.allclose(maps0.row_cdf[0].cpu().numpy(), np.arange(1, 25, dtype=F), rtol=1e-6)   # masses 1 / max(0, 1)
    # sampling still works on the uniform maps of an all-invalid query (sim = 0 everywhere)
    idx = pe.sample_correspondences(maps0, _dev(rng.random((1, 500, 2)).astype(F))).cpu().numpy()
    assert idx.min() >= 0 and (idx[..., 0] < 24).all() and (idx[..., 1:] < 256).all()
    assert len(np.unique(idx[0, :, 1])) > 100, "uniform soft-max: the draws spread over the map rows"


def test_ransac_single_retry_and_degenerate_sets():
    """num_retries = 1 skips the ratio test (:153-163); coincident correspondences give the identity rotation."""
    from snap_b200 import pose_estimation as pe, types
    i_xy = np.array([[0.0, 1.0], [2.0, 1.0], [0.5, 3.0]], F)
    idx = np.array([[[0, 10, 10], [1, 10, 20],        # pose 0: (0,1)->(2.1,2.1), (2,1)->(2.1,4.1): +90 degrees
                     [2, 5, 5], [2, 5, 5]]], np.int32)  # pose 1: the same correspondence twice
    poses = pe.transforms_from_correspondences(_dev(idx), _dev(i_xy), 2, 1, types.Grid2D((32, 32), 0.2)).cpu().numpy()[0]
    assert abs(poses[0, 0] - np.pi / 2) < 1e-5
    # t = mu_j - R mu_i with R = +90 degrees: mu_j = (2.1, 3.1), mu_i = (1, 1) -> R mu_i = (-1, 1)
    assert np.abs(poses[0, 1:] - np.array([3.1, 2.1], F)).max() < 1e-5
    assert poses[1, 0] == 0.0 and np.abs(poses[1, 1:] - (np.array([1.1, 1.1], F) - i_xy[2])).max() < 1e-6
