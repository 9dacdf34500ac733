"""`do_weighted_fusion=False` (streetview_encoder.py:262-267 without depth_mlp) through the product vs the oracle.

The product runs this branch on the kernels of the weighted branch (zero scale logits written by the GEMM engine, a zero
row appended to the fusion MLP's first kernel; `snap_b200/streetview_encoder.py`).  The identity is checked on the CPU
oracle in tests/test_golden.py::test_unweighted_lift_equals_weighted_lift_with_zero_logits, and the oracle branch is
pinned against the reference's own `StreetViewEncoder.__call__`.

First B200 run: round 2 (green).
"""
import numpy as np
import pytest
import torch

from util import F, rd_bf16, rel_l2, to_oracle_geometry

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


@pytest.mark.parametrize("V", [1, 3])
def test_unweighted_fusion_volume_and_fused_plane_vs_oracle(V):
    """Teacher-forced at the encoder output (the product's own finest FPN level is the oracle's `image_feature_pyr`):
    visibility and valid masks bit-exact; volume (unfused kernels) and plane (fused kernel) <= 5e-3 relative L2."""
    from oracle import bev_mapper as obm, grids as ogrids
    from snap_b200 import bev_mapper, configs, params, synthetic, types
    G, hw = 32, (96, 128)
    rng = np.random.default_rng(23)
    cfg = configs.bev_mapper(("streetview",))
    cfg.streetview_encoder.do_weighted_fusion = False
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    svp = p["streetview_encoder"]
    assert "proj_mlp" not in svp and svp["fusion_mlp"]["Dense_0"]["kernel"].shape == (256, 256)   # [mean | var]
    data = synthetic.make_tile(31, V, hw, G)
    grid = types.Grid2D((G, G), 0.2)
    mapper = bev_mapper.BEVMapper(cfg, grid)

    pred = mapper.apply({"params": p}, dict(data), debug=True)          # unfused kernels: the volume is materialised
    torch.cuda.synchronize()
    sv = pred["streetview"]
    assert "scores_images" not in sv and "feature_volume" in sv
    f_img = sv["image_feature_pyramid"].features[-1].float().cpu().numpy()[None]       # [1, V, hf, wf, 128]
    assert f_img.shape[1:] == (V, hw[0] // 4, hw[1] // 4, 128)
    vis = sv["debug"]["vis"][0].cpu().numpy().astype(bool)
    vol = sv["feature_volume"].features[0].float().cpu().numpy()
    vol_valid = sv["feature_volume"].valid[0].cpu().numpy().astype(bool)
    plane_u = sv["feature_plane"].features[0].float().cpu().numpy()
    pvalid_u = sv["feature_plane"].valid[0].cpu().numpy().astype(bool)

    pred_f = mapper.apply({"params": p}, dict(data))                    # fused single-kernel lift
    torch.cuda.synchronize()
    assert "feature_volume" not in pred_f["streetview"]
    plane_f = pred_f["streetview"]["feature_plane"].features[0].float().cpu().numpy()
    pvalid_f = pred_f["streetview"]["feature_plane"].valid[0].cpu().numpy().astype(bool)
    match_f = pred_f["bev_matching"].features[0].float().cpu().numpy()

    ocam, oT = to_oracle_geometry(data)
    ref = obm.bev_mapper_forward({"camera": ocam, "T_view2scene": oT}, p, ogrids.Grid2D((G, G), 0.2), rd=rd_bf16,
                                 return_volume=True, weighted=False,
                                 precomputed={"sv_features": f_img.astype(F), "sv_stride": (4.0, 4.0)})
    osv = ref["streetview"][0]
    assert np.array_equal(vis, osv["vis"])
    assert np.array_equal(vol_valid, osv["volume_valid"]) and 0.02 < vol_valid.mean() < 0.98
    assert np.array_equal(pvalid_u, osv["valid"]) and np.array_equal(pvalid_f, osv["valid"])
    e_vol = rel_l2(vol[vol_valid], osv["feature_volume"][osv["volume_valid"]])
    e_pu = rel_l2(plane_u, osv["feature_plane"])
    e_pf = rel_l2(plane_f, osv["feature_plane"])
    e_m = rel_l2(match_f, ref["bev_matching"]["features"][0])
    print(f"V={V}: rel_l2 volume {e_vol:.5f}, plane unfused {e_pu:.5f}, plane fused {e_pf:.5f}, matching {e_m:.5f}")
    assert not vol[~vol_valid].any() and not plane_f[~pvalid_f].any()
    assert e_vol < 1e-3 and e_pu < 1e-3 and e_pf < 1e-3 and e_m < 1e-3   # north_star bound; measured 8e-5 .. 2.5e-4


@pytest.mark.parametrize("V,layout", [(2, {}), (4, dict(spacing=0.5, same_side=True))])
def test_depth_mlp_residual_vs_oracle(V, layout):
    """`do_weighted_fusion=False` WITH the per-observation `depth_mlp` residual (streetview_encoder.py:263-267):
    f_proj += depth_mlp([f_proj, log10(clip(depth, 0.1, 100)), rays]) per (voxel, view), then plain mean / variance pooling.
    Product: snapb200_lift_observe -> the MLP on the tcgen05 GEMM engine (N * V rows) -> snapb200_lift_pool_observations ->
    fusion MLP -> vertical max.  Teacher-forced at the encoder output; the oracle branch is pinned against the reference's
    own `StreetViewEncoder.__call__` (tests/golden/sve_call_plain_depthmlp_*.npz)."""
    from oracle import bev_mapper as obm, grids as ogrids
    from snap_b200 import bev_mapper, configs, params, synthetic, types
    from util import record_parity
    G, hw = 32, (96, 128)
    rng = np.random.default_rng(29)
    cfg = configs.bev_mapper(("streetview",))
    cfg.streetview_encoder.do_weighted_fusion = False
    cfg.streetview_encoder.depth_mlp = configs.mlp()
    cfg.streetview_encoder.depth_mlp.layers = (64, 128)
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg)))
    svp = p["streetview_encoder"]
    assert svp["depth_mlp"]["Dense_0"]["kernel"].shape == (132, 64) and "proj_mlp" not in svp
    data = synthetic.make_tile(33, V, hw, G, **layout)
    mapper = bev_mapper.BEVMapper(cfg, types.Grid2D((G, G), 0.2))
    pred = mapper.apply({"params": p}, dict(data), debug=True)
    torch.cuda.synchronize()
    sv = pred["streetview"]
    assert "feature_volume" in sv
    f_img = sv["image_feature_pyramid"].features[-1].float().cpu().numpy()[None]
    vis = sv["debug"]["vis"][0].cpu().numpy().astype(bool)
    vol = sv["feature_volume"].features[0].float().cpu().numpy()
    vol_valid = sv["feature_volume"].valid[0].cpu().numpy().astype(bool)
    plane = sv["feature_plane"].features[0].float().cpu().numpy()
    ocam, oT = to_oracle_geometry(data)
    ref = obm.bev_mapper_forward({"camera": ocam, "T_view2scene": oT}, p, ogrids.Grid2D((G, G), 0.2), rd=rd_bf16,
                                 return_volume=True, weighted=False,
                                 precomputed={"sv_features": f_img.astype(F), "sv_stride": (4.0, 4.0)})
    osv = ref["streetview"][0]
    assert np.array_equal(vis, osv["vis"]) and np.array_equal(vol_valid, osv["volume_valid"])
    multi = float((osv["vis"].sum(-1) > 1).sum()) / max(1, int(osv["volume_valid"].sum()))
    e_vol = rel_l2(vol[vol_valid], osv["feature_volume"][osv["volume_valid"]])
    e_pl = rel_l2(plane, osv["feature_plane"])
    e_m = rel_l2(pred["bev_matching"].features[0].float().cpu().numpy(), ref["bev_matching"]["features"][0])
    print(f"depth_mlp V={V}: valid frac {vol_valid.mean():.3f} ({multi:.0%} multi-view), rel_l2 volume {e_vol:.5f}, plane {e_pl:.5f}, "
          f"matching {e_m:.5f}")
    record_parity("depth_mlp branch (un-weighted fusion)", f"volume rel-L2 vs oracle, V={V}", e_vol, 2e-3)
    if layout:
        assert multi > 0.3
    # log10f / the ray normalisation differ from NumPy by an ulp before their bf16 rounding and feed a 2-layer MLP
    assert e_vol < 2e-3 and e_pl < 2e-3 and e_m < 2e-3
    assert not vol[~vol_valid].any()
    # the same module without the residual gives a different volume (the branch is live)
    with pytest.raises(NotImplementedError):
        mapper.streetview_encoder.apply({"params": svp}, {**data, "xyz_grid": mapper.build_xyz_grid(dict(data))}, fused=True)
