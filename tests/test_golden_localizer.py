"""The localizer oracle (oracle/pose_estimation.py) and the host-side geometry of snap_b200.bev_localizer against golden
fixtures produced by the REFERENCE'S OWN snap/models/pose_estimation.py and snap/models/bev_localizer.py executed
under the NumPy stand-in for jax (tests/golden/make_golden_localizer.py)."""
import os

import numpy as np

from oracle import grids, pose_estimation as ope

F = np.float32
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name + ".npz")))


def ang_close(a, b, tol):
    d = np.abs((np.asarray(a, np.float64) - np.asarray(b, np.float64) + np.pi) % (2 * np.pi) - np.pi)
    assert d.max() <= tol, d.max()


def test_kabsch():
    d = load("loc_kabsch")
    for c in range(5):
        i, j, angle, t, valid, rssd = (d[f"{k}_{c}"] for k in range(6))
        a, tt, v, r = ope.kabsch_algorithm_2d(i, j)
        ang_close(a, angle, 1e-5)
        assert np.abs(tt - t).max() <= 2e-5 * (np.abs(t).max() + 1)
        assert bool(v) == bool(valid)
        assert abs(float(r) - float(rssd)) <= 2e-3 * (abs(float(rssd)) + 1)
        if len(i) == 2:   # closed form used by the CUDA kernel (csrc/localizer.cu: loc_ransac_poses_kernel)
            A, B = i - i.mean(0), j - j.mean(0)
            cov = np.einsum("ji,jk->ik", A, B)
            ang_close(np.arctan2(cov[1, 0] - cov[0, 1], cov[0, 0] + cov[1, 1]), angle, 1e-5)


def test_pose_scoring_and_refinement():
    d = load("loc_scoring")
    H, W = d["scores_all"].shape[1:]
    grid = grids.Grid2D((H, W), 0.2)
    for mask in (0, 1):
        s = ope.pose_scoring_many(d["angle"], d["t"], d["scores_all"], d["i_xy"], d["valid_points"], d["valid_j"], grid,
                                  bool(mask))
        assert np.abs(s - d[f"scores_mask{mask}"]).max() <= 1e-5
        a, t, vol = ope.grid_refinement(d["init_angle"], d["init_t"], d["scores_all"], d["i_xy"], d["valid_points"],
                                        d["valid_j"], grid, bool(mask))
        assert vol.shape == d[f"refine_volume_mask{mask}"].shape == (41, 41, 41)
        assert np.abs(vol - d[f"refine_volume_mask{mask}"]).max() <= 2e-5
        ang_close(a, d[f"refined_angle_mask{mask}"], 1e-6)
        assert np.abs(t - d[f"refined_t_mask{mask}"]).max() <= 1e-5


def test_ransac_downstream_of_the_draw():
    d = load("loc_ransac")
    N, H, W = d["prob"].shape
    grid = grids.Grid2D((H, W), 0.5)
    for tag, (num_poses, retries) in {"r1": (12, 1), "r4": (12, 4)}.items():
        idx = np.stack(np.unravel_index(d[f"flat_{tag}"], (N, H, W)), -1)
        a, t = ope.sample_transforms_ransac(idx, d["i_xy_p"], num_poses, retries, grid)
        ang_close(a, d[f"angle_{tag}"], 2e-5)
        assert np.abs(t - d[f"t_{tag}"]).max() <= 1e-4


def test_inverse_cdf_sampler_equals_flat_searchsorted():
    """the nested (point, cell) inverse CDF of the oracle / the CUDA kernel is the flat jax.random.choice draw when
    both uniforms describe the same flat position"""
    d = load("loc_ransac")
    p = d["prob"].astype(np.float64)
    N, H, W = p.shape
    rng = np.random.default_rng(3)
    u = rng.random((500, 2)).astype(F)
    idx = ope.sample_correspondences_inverse_cdf(d["prob"], u)
    # distribution check: every drawn cell has positive probability and the marginal over points follows row masses
    assert (p[idx[:, 0], idx[:, 1], idx[:, 2]] > 0).all()
    row_mass = p.reshape(N, -1).sum(-1)
    cdf = np.cumsum(row_mass)
    n_flat = np.searchsorted(cdf, cdf[-1] * (1.0 - u[:, 0].astype(np.float64)), side="left")
    assert np.array_equal(idx[:, 0], n_flat)


def test_frustum_grid_and_loss():
    from snap_b200 import bev_localizer as bl, types
    d = load("loc_localizer")
    g, gp, q = bl.build_query_frustum_grid(0.2, 16.0, True, 72.0)
    assert tuple(g.extent) == tuple(d["extent"]) and np.array_equal(gp, d["grid_p_view"])
    # the stand-in evaluates in float64 (JAX: fp32): same points to fp32 round-off, same field-of-view selection
    assert q.shape == d["q_xy_p"].shape and np.abs(q - d["q_xy_p"]).max() <= 2e-6
    g2, _, q2 = bl.build_query_frustum_grid(0.5, 8.0, False, None)
    assert tuple(g2.extent) == tuple(d["extent2"]) and np.abs(q2 - d["q_xy_p2"]).max() <= 2e-6
    og, ogp, oq = ope.build_query_frustum_grid(0.5, 8.0)
    assert np.abs(oq - d["q_xy_p2"]).max() <= 2e-6
    gt = bl.transform2d_from_transform3d(types.Transform3D(R=d["gt_R"], t=d["gt_t"]))
    assert np.abs(gt[:, 0] - d["gt_angle"]).max() <= 1e-6 and np.array_equal(gt[:, 1:], d["gt_t"][:, :2])
    for b in range(len(d["scores"])):
        nll, m, dr_s, dt_s = ope.loss_metrics(d["scores"][b], d["samples_angle"][b], d["samples_t"][b], d["best_angle"][b],
                                              d["best_t"][b], gt[b, 0], gt[b, 1:], None)
        assert abs(float(nll) - float(d["nll"][b])) <= 1e-5 * (1 + abs(float(d["nll"][b])))
        assert abs(m["loc/err_max_position"] - d["err_pos"][b]) <= 1e-5
        assert abs(m["loc/err_max_rotation"] - d["err_rot"][b]) <= 1e-3
        assert m["loc/recall_top1"] == bool(d["top1"][b])
        for k, key in enumerate(["loc/recall_samples_0.5m_1", "loc/recall_samples_1m_2", "loc/recall_samples_2m_4"]):
            assert abs(m[key] - d[f"rec{k}"][b]) <= 1e-6
        assert d["rec2"][b] > 0   # the fixture plants near-ground-truth samples


def test_matching_block_vs_reference_bevlocalizer_call():
    """oracle point_similarities + pose_scoring_many against the reference's OWN BEVLocalizer.__call__ (bev_localizer.py:131-218)
    run under the stand-in on given BEV planes (tests/golden/make_golden_localizer_call.py): similarity normalisation
    (1/num_valid or the masked soft-max of the query confidences), temperature, scoring with and without the out-of-bounds
    mask, ground truth prepended, arg-max over the sampled poses."""
    for tag in ("plain", "conf_mask"):
        d = load("loc_call_" + tag)
        B, N = d["vq"].shape[:2]
        H, W = d["vm"].shape[1:]
        grid = grids.Grid2D((H, W), float(d["cell"]))
        gt_angle, gt_t = ope.transform2d_from_transform3d(d["gt_R"], d["gt_t"])
        for b in range(B):
            assert abs(d["samples_angle"][b, 0] - gt_angle[b]) < 1e-6 and np.array_equal(d["samples_t"][b, 0], gt_t[b])
            sim, prob = ope.point_similarities(d["fq"][b, :, 0], d["vq"][b, :, 0], d["fm"][b], float(d["temperature"]), True,
                                               d["conf"][b, :, 0] if bool(d["add_conf"]) else None)
            sc = ope.pose_scoring_many(d["samples_angle"][b], d["samples_t"][b], sim, d["q_xy_p"][:, 0], d["vq"][b, :, 0],
                                       d["vm"][b], grid, bool(d["mask_oob"]))
            assert np.abs(sc - d["scores_poses"][b]).max() <= 2e-5 * (1 + np.abs(d["scores_poses"][b]).max()), (tag, b)
            k = int(np.argmax(sc[1:]))
            assert k == int(d["best_index"][b])
            assert abs(d["samples_angle"][b, 1 + k] - d["best_angle"][b]) < 1e-7 and np.array_equal(d["samples_t"][b, 1 + k], d["best_t"][b])
            # prob_points: every point's map sums to its weight (1 / num_valid, or its confidence soft-max weight)
            w = prob.reshape(N, -1).sum(-1)
            if bool(d["add_conf"]):
                assert abs(w.sum() - 1) < 1e-5 and not w[~d["vq"][b, :, 0]].any()
            else:
                assert np.allclose(w, 1.0 / max(int(d["vq"][b].sum()), 1), rtol=1e-5)


def test_loss_with_removed_accurate_poses_and_dense_plane_recovery():
    """threshold_remove_accurate_poses (bev_localizer.py:254-258) in the oracle loss, and the product's host-side
    recover_dense_feature_plane (:111-129), against the reference's own methods run under the stand-in."""
    import torch
    from snap_b200 import bev_localizer as bl, configs, types
    d = load("loc_localizer")
    gt = bl.transform2d_from_transform3d(types.Transform3D(R=d["gt_R"], t=d["gt_t"]))
    changed = 0
    for b in range(len(d["scores"])):
        nll, _, dr_s, dt_s = ope.loss_metrics(d["scores"][b], d["samples_angle"][b], d["samples_t"][b], d["best_angle"][b],
                                              d["best_t"][b], gt[b, 0], gt[b, 1:], (1.5, 0.6))
        assert abs(float(nll) - float(d["nll_removed"][b])) <= 1e-5 * (1 + abs(float(d["nll_removed"][b])))
        changed += abs(float(d["nll_removed"][b]) - float(d["nll"][b])) > 1e-6
        assert ((dr_s[1:] < 1.5) & (dt_s[1:] < 0.6)).any(), "the fixture plants samples inside the removal thresholds"
    assert changed > 0
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov, cfg.num_pose_samples = True, 8
    loc = bl.BEVLocalizer(cfg, None, types.Grid2D((32, 32), 0.2))
    dense = loc.recover_dense_feature_plane(types.FeaturePlane(torch.from_numpy(d["sparse_features"]),
                                                               torch.from_numpy(d["sparse_valid"].astype(np.uint8))))
    assert np.array_equal(dense.valid.numpy().astype(bool), d["dense_valid"].astype(bool))
    assert np.array_equal(dense.features.numpy(), d["dense_features"].astype(F))
