"""Golden fixture for the matching block of BEVLocalizer: the reference's OWN `BEVLocalizer.__call__`
(bev_localizer.py:131-218) executed under the NumPy stand-in for jax on a stand-in `self` whose `bev_mapper` returns
given BEV planes.  `jax.random.choice` is the documented inverse-CDF draw on NumPy generators (the fixture stores the
sampled poses, so everything downstream of the draw is pinned).  Run in the build container only:

    python tests/golden/make_golden_localizer_call.py     # writes tests/golden/loc_call_*.npz
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
from snap.models import bev_localizer as bl, types as rtypes  # noqa: E402

F = np.float32


class Cfg(dict):
    __getattr__ = dict.__getitem__


def unit(x):
    return (x / np.linalg.norm(x, axis=-1, keepdims=True)).astype(F)


def case(tag, add_conf, mask_oob, seed):
    rng = np.random.default_rng(seed)
    B, N, H, W, D = 2, 30, 9, 11, 6
    grid = grids.Grid2D((H, W), 0.5)
    q_xy_p = ((rng.random((N, 1, 2)) - [0.5, 0.0]) * [3.0, 2.5]).astype(F)
    fm = unit(rng.standard_normal((B, H, W, D)))
    vm = rng.random((B, H, W)) < 0.85
    fm = fm * vm[..., None]
    vq = rng.random((B, N, 1)) < 0.7
    fq = unit(rng.standard_normal((B, N, 1, D))) * vq[..., None]
    conf = rng.standard_normal((B, N, 1)).astype(F)
    planes = {"map": {"bev_matching": rtypes.FeaturePlane(features=fm, valid=vm)},
              "query": {"bev_matching": rtypes.FeaturePlane(features=fq, valid=vq), "bev_confidence": conf}}
    cfg = Cfg(add_confidence_query=add_conf, clip_negative_scores=True, add_temperature=True, num_pose_samples=12,
              num_pose_sampling_retries=2, mask_score_out_of_bounds=mask_oob, do_grid_refinement=False)
    fake = pytypes.SimpleNamespace(
        config=cfg, grid_map=grid, q_xy_p=q_xy_p, temperature=np.asarray(1.3, F), bev_mapper_query=None,
        bev_mapper=lambda data, train, debug, is_query=False: planes["query" if is_query else "map"],
        make_rng=lambda name: np.random.default_rng(seed + 100))
    a = rng.uniform(-3, 3, B)
    T3 = geometry.Transform3D(R=np.stack([[[np.cos(x), -np.sin(x), 0], [np.sin(x), np.cos(x), 0], [0, 0, 1]] for x in a]).astype(F),
                              t=(rng.random((B, 3)) * [4, 5, 0]).astype(F))
    data = {"map": {}, "query": {"images": np.zeros((B, 1))}, "T_query2map": T3}
    pred = bl.BEVLocalizer.__call__(fake, data, False, False)
    out = dict(q_xy_p=q_xy_p, fm=fm, vm=vm, fq=fq, vq=vq, conf=conf, temperature=np.asarray(1.3, F), cell=np.asarray(0.5),
               gt_R=T3.R, gt_t=T3.t, samples_angle=pred["map_t_query_samples"].angle, samples_t=pred["map_t_query_samples"].t,
               scores_poses=pred["scores_poses"], best_index=pred["best_index"], best_angle=pred["map_t_query"].angle,
               best_t=pred["map_t_query"].t, add_conf=np.asarray(add_conf), mask_oob=np.asarray(mask_oob))
    np.savez_compressed(os.path.join(HERE, f"loc_call_{tag}.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(tag, {k: np.asarray(v).shape for k, v in out.items() if k in ("samples_angle", "scores_poses", "best_index")},
          np.asarray(pred["scores_poses"])[0, :4])


case("plain", False, False, 11)
case("conf_mask", True, True, 12)
