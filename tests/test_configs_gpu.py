"""BASELINE.json configs 3 and 4 at their full sizes, through size-independent properties (an oracle run at these
sizes takes minutes on the CPU): fused == unfused lift, unit-norm matching features, known-pose vote peak."""
import numpy as np
import pytest
import torch

from util import F, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _subset_views(data, views):
    """Batch dict restricted to a subset of its views (same scene, same cameras)."""
    from snap_b200 import types
    cam, T = data["camera"], data["T_view2scene"]
    out = {"images": np.ascontiguousarray(data["images"][:, views]),
           "camera": types.Camera(wh=cam.wh[:, views].copy(), f=cam.f[:, views].copy(), c=cam.c[:, views].copy()),
           "T_view2scene": types.Transform3D(R=T.R[:, views].copy(), t=T.t[:, views].copy())}
    return out


def test_config3_streetview_aerial_g256():
    """configs[2]: StreetView + aerial fusion on a 256 x 256 grid (the reference's R50 encoder)."""
    from snap_b200 import bev_mapper, configs, params, synthetic, types
    G, V, hw = 256, 4, (480, 640)
    rng = np.random.default_rng(5)
    cfg = configs.bev_mapper(("streetview", "aerial"))
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    data = synthetic.make_tile(31, V, hw, G, aerial=True)
    grid = types.Grid2D((G, G), 0.2)
    fused = bev_mapper.BEVMapper(cfg, grid).apply({"params": p}, dict(data))
    f_feat, f_valid = fused["bev_features"].features.float().cpu().numpy(), fused["bev_features"].valid.cpu().numpy()
    f_sv = fused["streetview"]["feature_plane"]
    f_sv_feat, f_sv_valid = f_sv.features.float().cpu().numpy(), f_sv.valid.cpu().numpy().astype(bool)
    f_match = fused["bev_matching"].features.float().cpu().numpy()
    unf = bev_mapper.BEVMapper(cfg, grid, fused_lift=False).apply({"params": p}, dict(data))
    torch.cuda.synchronize()
    assert "feature_volume" in unf["streetview"] and "feature_volume" not in fused["streetview"]
    u_sv = unf["streetview"]["feature_plane"]
    assert np.array_equal(f_sv_valid, u_sv.valid.cpu().numpy().astype(bool)), "street-view valid plane: fused != unfused"
    assert 0.02 < f_sv_valid.mean() < 0.9
    a, b = f_sv_feat, u_sv.features.float().cpu().numpy()
    print(f"cfg3: street-view valid cells {int(f_sv_valid.sum())}, fused vs unfused differing elements {(a != b).mean():.5%}")
    assert (a != b).mean() < 2e-3 and rel_l2(a, b) < 1e-3
    # aerial covers every cell, so the fused map is valid everywhere and >= the street-view plane where that is valid
    assert f_valid.all()
    assert (f_feat[0][f_sv_valid[0]] >= f_sv_feat[0][f_sv_valid[0]]).all(), "modality max must dominate its inputs"
    nrm = np.linalg.norm(f_match, axis=-1)
    assert np.abs(nrm - 1).max() < 2e-2, "matching features are L2-normalised (bf16 rounding of 32 components)"
    assert np.array_equal(unf["bev_features"].valid.cpu().numpy(), f_valid)


def test_config4_localization_pipeline_g128_r36():
    """configs[3] per example: map BEV (4 views), query BEV (1 of those views), exhaustive voting with 36 rotations.
    Cells that only the query view sees have identical pooled features in both maps, so the vote must peak at the
    identity pose (rotation 0, zero shift = index G-1)."""
    from snap_b200 import bev_mapper, configs, params, pose_exhaustive_voting as pv, synthetic, types
    G, R, hw = 128, 36, (480, 640)
    rng = np.random.default_rng(6)
    cfg = configs.bev_mapper(("streetview",))
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    data = synthetic.make_tile(41, 4, hw, G)
    grid = types.Grid2D((G, G), 0.2)
    mapper = bev_mapper.BEVMapper(cfg, grid)
    # one voxel grid for both maps: z_offset = median camera height - scene_z_offset (bev_mapper.py:171-174)
    z_off = (np.median(data["T_view2scene"].t[..., -1].astype(F), axis=-1).astype(F) - F(4.0)).astype(F)
    data["z_offset"] = z_off
    pm = mapper.apply({"params": p}, dict(data))["bev_matching"]
    pm = types.FeaturePlane(pm.features.clone(), pm.valid.clone())      # the mapper reuses its output buffers
    qdata = _subset_views(data, [1])
    qdata["z_offset"] = z_off
    pq = mapper.apply({"params": p}, qdata, is_query=True)["bev_matching"]
    scores = pv.exhaustive_pose_voting(pq, pm, R, grid)
    torch.cuda.synchronize()
    s = scores[0].cpu().numpy()
    assert s.shape == (R, 2 * G - 1, 2 * G - 1)
    fin = np.isfinite(s)
    assert fin.any() and np.isneginf(s[~fin]).all(), "non-finite scores are exactly the -inf min-overlap mask"
    qv, mv = pq.valid[0].cpu().numpy().astype(bool), pm.valid[0].cpu().numpy().astype(bool)
    assert qv.sum() > 500 and (qv & ~mv).sum() == 0, "every cell the query sees is also seen by the map"
    k = np.unravel_index(np.argmax(np.where(fin, s, -np.inf)), s.shape)
    print(f"cfg4: query cells {int(qv.sum())}, map cells {int(mv.sum())}, peak {k} score {s[k]:.4f}")
    assert tuple(int(x) for x in k) == (0, G - 1, G - 1)


def test_cfg2_full_size_teacher_forced_parity_vs_oracle():
    """BASELINE configs[1] at FULL size (4 x 480x640 views, 128 x 128 x 60 = 983,040 voxels), the benchmarked configuration,
    compared NUMERICALLY with the oracle: the valid plane bit-exact, and -- the oracle (bf16 emulation) being fed the GPU's
    own encoder features, which removes the chaos of a random-init 50-layer bf16 ResNet -- proj MLP -> lift -> fusion MLP ->
    vertical max -> matching head to <= 1e-3 relative L2 (north_star).  The same check runs inside bench.py on the
    cpu_baseline tile and is printed as the `parity` object of the JSON line."""
    import os
    import bench
    from util import record_parity
    from snap_b200 import bev_mapper, configs, params, types
    cfg = configs.bev_mapper(("streetview",))
    p = params.round_to_bf16(params.init_bev_mapper(np.random.default_rng(7), cfg))
    mapper = bev_mapper.BEVMapper(cfg, types.Grid2D((bench.G, bench.G), 0.2))
    parity, _ = bench.cfg2_parity(mapper, p, 99, os.cpu_count() or 1, torch.device("cuda"), free_running=False)
    tf = parity["teacher_forced"]
    print(f"cfg2 full size: valid_equal {parity['valid_equal']} ({parity['valid_cells']} valid cells), teacher-forced rel_l2 "
          f"bev_features {tf['rel_l2_bev_features']:.2e}, bev_matching {tf['rel_l2_bev_matching']:.2e}")
    record_parity("cfg2 full size (teacher-forced lift + head)", "bev_features rel-L2 vs oracle", tf["rel_l2_bev_features"], 1e-3)
    record_parity("cfg2 full size (teacher-forced lift + head)", "bev_matching rel-L2 vs oracle", tf["rel_l2_bev_matching"], 1e-3)
    assert parity["valid_equal"] and parity["valid_cells"] > 5000
    assert tf["rel_l2_bev_features"] <= 1e-3 and tf["rel_l2_bev_matching"] <= 1e-3
