"""Golden fixtures for the sampling localizer: the reference's OWN `snap/models/pose_estimation.py`
and the pure parts of `snap/models/bev_localizer.py`, executed under the NumPy/SciPy stand-in for jax
(tests/golden/jaxshim).  Run in the build container only:

    python tests/golden/make_golden_localizer.py     # writes tests/golden/loc_*.npz

`jax.random.choice` is replaced by its documented inverse-CDF algorithm on a NumPy generator (the threefry stream
itself cannot be reproduced without JAX): the fixture stores the drawn indices, so everything downstream is pinned.
"""
import os
import sys
import types as pytypes

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")

import jaxshim  # noqa: E402

jaxshim.install(REF)

from snap.utils import geometry, grids  # noqa: E402
from snap.models import pose_estimation as pe  # noqa: E402

F = np.float32
rng = np.random.default_rng(20241017)
out = {}

# ---- kabsch_algorithm_2d (:103-123) --------------------------------------------------------------------
cases = []
for n in (2, 2, 2, 5, 9):
    j = (rng.standard_normal((n, 2)) * 4).astype(F)
    a = rng.uniform(-np.pi, np.pi)
    Rm = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    i = (j @ Rm.T + rng.standard_normal(2) * 3 + rng.standard_normal((n, 2)) * 0.05).astype(F)
    T, valid, rssd = pe.kabsch_algorithm_2d(i, j)
    cases.append((i, j, T.angle, T.t, valid, rssd))
out["loc_kabsch"] = {f"{k}_{c}": np.asarray(v[k]) for c, v in enumerate(cases)
                     for k in range(6)}

# ---- pose_scoring (:65-85) and grid_refinement (:168-203) ------------------------------------------------
N, H, W = 3, 9, 11
grid = grids.Grid2D((H, W), 0.2)
scores_all = rng.random((N, H, W)).astype(F)
i_xy = ((rng.random((N, 2)) - 0.5) * 1.2).astype(F)
valid_points = np.array([True, True, False])
valid_j = rng.random((H, W)) > 0.15
poses = geometry.Transform2D(angle=rng.uniform(-3, 3, 40).astype(F),
                             t=(rng.random((40, 2)) * [H * 0.2, W * 0.2] * 1.4 - 0.2).astype(F))
d = dict(scores_all=scores_all, i_xy=i_xy, valid_points=valid_points, valid_j=valid_j, angle=poses.angle, t=poses.t)
for mask in (False, True):
    d[f"scores_mask{int(mask)}"] = pe.pose_scoring_many(poses, scores_all, i_xy, valid_points, valid_j, grid, mask)
init = geometry.Transform2D(angle=np.asarray(0.4, F), t=np.asarray([0.9, 1.1], F))
for mask in (() if "--skip-refinement" in sys.argv else (False, True)):   # 68,921 poses through the stand-in: ~1 min each
    ref, vol = pe.grid_refinement(init, scores_all, i_xy, valid_points, valid_j, grid, mask)
    d[f"refined_angle_mask{int(mask)}"], d[f"refined_t_mask{int(mask)}"] = ref.angle, ref.t
    d[f"refine_volume_mask{int(mask)}"] = vol
d["init_angle"], d["init_t"] = init.angle, init.t
out["loc_scoring"] = d

# ---- sample_transforms_ransac (:126-165) -----------------------------------------------------------------
N, H, W = 6, 7, 8
grid = grids.Grid2D((H, W), 0.5)
prob = rng.random((N, H, W)).astype(F) ** 4
prob /= prob.sum()
i_xy_p = ((rng.random((N, 2)) - 0.5) * 6).astype(F)
d = dict(prob=prob, i_xy_p=i_xy_p)
drawn = []
orig_choice = jaxshim.random_choice


def recording_choice(r, a, shape=(), replace=True, p=None):
    idx = orig_choice(r, a, shape, replace, p)
    drawn.append(idx)
    return idx


sys.modules["jax"].random.choice = recording_choice
for tag, (num_poses, retries) in {"r1": (12, 1), "r4": (12, 4)}.items():
    T = pe.sample_transforms_ransac(np.random.default_rng(5), prob, i_xy_p, num_poses, retries, grid)
    d[f"flat_{tag}"], d[f"angle_{tag}"], d[f"t_{tag}"] = drawn[-1], T.angle, T.t
out["loc_ransac"] = d

# ---- bev_localizer: frustum grid, loss / metrics ---------------------------------------------------------------
try:
    from snap.models import bev_localizer as bl
    g, gp, q = bl.build_query_frustum_grid(0.2, 16.0, True, 72.0)
    g2, gp2, q2 = bl.build_query_frustum_grid(0.5, 8.0, False, None)
    d = dict(extent=np.asarray(g.extent), grid_p_view=gp, q_xy_p=q, extent2=np.asarray(g2.extent), q_xy_p2=q2)
    # loss_metrics_function (:244-278) without threshold_remove_accurate_poses
    B, P1 = 3, 50
    scores = (rng.standard_normal((B, P1)) * 3).astype(F)
    samples = geometry.Transform2D(angle=rng.uniform(-3.2, 3.2, (B, P1)).astype(F),
                                   t=(rng.standard_normal((B, P1, 2)) * 3).astype(F))
    Rz = lambda a: np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], F)
    gt_a = rng.uniform(-3, 3, B)
    T3 = geometry.Transform3D(R=np.stack([Rz(a) for a in gt_a]), t=(rng.standard_normal((B, 3)) * 2).astype(F))
    gt2 = geometry.Transform2D.from_Transform3D(T3)
    # plant near-GT samples so that the recalls are not all zero
    samples.angle[:, 1:6] = gt2.angle[:, None] + rng.uniform(-0.03, 0.03, (B, 5)).astype(F)
    samples.t[:, 1:6] = gt2.t[:, None] + rng.uniform(-0.7, 0.7, (B, 5, 2)).astype(F)
    samples.angle[:, 0], samples.t[:, 0] = gt2.angle, gt2.t
    best = samples[:, 3]
    self = pytypes.SimpleNamespace(config=pytypes.SimpleNamespace(threshold_remove_accurate_poses=None,
                                                                  add_temperature=False))
    pred = {"scores_poses": scores, "map_t_query_samples": samples, "map_t_query": best}
    losses, metrics = bl.BEVLocalizerModel.loss_metrics_function(self, pred, {"T_query2map": T3})
    d.update(scores=scores, samples_angle=samples.angle, samples_t=samples.t, gt_R=T3.R, gt_t=T3.t,
             gt_angle=gt2.angle, best_angle=best.angle, best_t=best.t, nll=losses["total"],
             err_pos=metrics["loc/err_max_position"], err_rot=metrics["loc/err_max_rotation"],
             top1=metrics["loc/recall_top1"],
             rec0=metrics["loc/recall_samples_0.5m_1°"], rec1=metrics["loc/recall_samples_1m_2°"],
             rec2=metrics["loc/recall_samples_2m_4°"])
    # the same with threshold_remove_accurate_poses (:254-258: accurate samples are removed from the soft-max, never index 0)
    self_t = pytypes.SimpleNamespace(config=pytypes.SimpleNamespace(threshold_remove_accurate_poses=(1.5, 0.6), add_temperature=False))
    losses_t, _ = bl.BEVLocalizerModel.loss_metrics_function(self_t, pred, {"T_query2map": T3})
    d["nll_removed"] = losses_t["total"]
    # recover_dense_feature_plane (:111-129) on the field-of-view points
    fake_loc = pytypes.SimpleNamespace(grid_query=g, qgrid_p_q=gp, q_xy_p=q)
    from snap.models import types as rtypes
    sparse = rtypes.FeaturePlane(features=rng.standard_normal((len(q), 1, 3)).astype(F), valid=rng.random((len(q), 1)) < 0.7)
    dense = bl.BEVLocalizer.recover_dense_feature_plane(fake_loc, sparse)
    d.update(sparse_features=sparse.features, sparse_valid=sparse.valid, dense_features=dense.features, dense_valid=dense.valid)
    out["loc_localizer"] = d
except Exception as e:  # pragma: no cover
    import traceback
    traceback.print_exc()
    print("bev_localizer could not be executed under the stand-in:", e)

if "--skip-refinement" in sys.argv:
    del out["loc_scoring"]
for name, d in out.items():
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **{k: np.asarray(v) for k, v in d.items()})
    print(name, {k: np.asarray(v).shape for k, v in d.items()})
