"""The oracle (oracle/) against golden fixtures produced by the REFERENCE'S OWN source executed under a NumPy/SciPy
stand-in for jax (tests/golden/make_golden.py, tests/golden/jaxshim).  Boolean / integer outputs must match exactly;
floats to fp32 round-off (the stand-in evaluates some expressions in float64)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometry, grids, image_encoder as oie, layers, pose_exhaustive_voting as opv, resnet as ores
from oracle import streetview_encoder as osv

F = np.float32
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(G, name + ".npz")))


def close(a, b, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True)
    scale = np.abs(b[fin]).max() + 1e-12 if fin.any() else 1.0
    assert np.abs(a[fin] - b[fin]).max() <= tol * scale, np.abs(a[fin] - b[fin]).max() / scale


def _setup():
    g = load("geometry")
    T = geometry.Transform3D(R=g["R"], t=g["t"])
    cam = geometry.Camera(wh=g["wh"], f=g["f"], c=g["c"])
    return g, T, cam


def test_interpolate_nd():
    d = load("interp")
    v, ok = grids.interpolate_nd(d["arr"], d["pts"])
    close(v, d["val"]); assert np.array_equal(ok, d["valid"])
    _, ok2 = grids.interpolate_nd(d["arr"], d["pts"], d["mask"])
    assert np.array_equal(ok2, d["valid_masked"])


def test_geometry_and_projection():
    g, T, cam = _setup()
    Ti = T.inv
    close(Ti.R, g["Rinv"], 1e-7); close(Ti.t, g["tinv"], 1e-6)
    close(cam.scale(F([0.25, 0.25])).f, g["cam_scaled_f"], 1e-7)
    p = load("project")
    p2d, vis, depth, rays = osv.project_points_to_views(T, cam, g["pts"])
    # visibility may only differ where a point sits within fp32 round-off of an image border / eps plane
    margin = np.minimum.reduce([np.abs(p["p2d"][..., 0]), np.abs(p["p2d"][..., 1]),
                                np.abs(p["p2d"][..., 0] - cam.wh[None, :, 1]), np.abs(p["p2d"][..., 1] - cam.wh[None, :, 0])])
    strict = margin > 1e-3
    assert np.array_equal(vis[strict], p["vis"][strict]) and (vis != p["vis"]).mean() < 1e-2
    sel = p["vis"] & vis
    close(p2d[sel], p["p2d"][sel], 1e-5); close(depth, p["depth"], 1e-5); close(rays, p["rays"], 1e-5)
    pf = load("project_fisheye")
    fcam = geometry.FisheyeCamera(wh=g["wh"], f=g["f"], c=g["c"], k_radial=g["k_radial"], max_fov=g["max_fov"])
    p2df, visf, _, _ = osv.project_points_to_views(T, fcam, g["pts"])
    assert (visf != pf["vis"]).mean() < 1e-2
    sel = pf["vis"] & visf
    close(p2df[sel], pf["p2d"][sel], 1e-4)


def test_lift_pieces():
    g, T, cam = _setup()
    d = load("interp_views_all")
    out = osv.interpolate_views_all(d["fimg"][0], d["p2d"])
    close(out, d["out"][0])
    ds = load("depth_score")
    close(osv.interpolate_depth_score(ds["scales"][0], ds["depth"]), ds["out"][0])
    for tag, use_scores, minmax in (("weighted", True, False), ("plain", False, False), ("minmax", True, True)):
        pw = load("pool_" + tag)
        base = load("pool_weighted")
        st, va = osv.pool_multiview_features(base["feats"][0], base["valid"], ds["out"][0] if use_scores else None, minmax, True)
        close(st, pw["stats"][0], 5e-5)
        if "valid_any" in pw:
            assert np.array_equal(va, pw["valid_any"][0])
    vs = load("view_selection")
    idx, md = osv.view_selection(g["pts"], T, load("pool_weighted")["valid"], 2)
    assert np.array_equal(idx, vs["idx"][0]); close(md, vs["min_dist"][0])
    sl = load("interp_views_selective")
    close(osv.interpolate_views_selective(d["fimg"][0], sl["p2d"][0], sl["idx"][0]), sl["out"][0])


def test_layers_resnet_pad():
    d = load("normalize")
    close(layers.normalize(d["x"]), d["out"], 1e-6)
    s = load("standardize")
    close(ores.standardize(torch.from_numpy(s["w"]), [0, 1, 2], 1e-10).numpy(), s["out"], 1e-5)
    close(ores.standardize(torch.from_numpy(s["w"]).reshape(1, 3, 3, 5, 4), [1, 2, 4], 1e-5).numpy(), s["gn"], 1e-5)
    p = load("pad")
    for key, stride in (("p8", 8), ("p32", 32)):
        assert np.array_equal(oie.pad_to_multiple(torch.from_numpy(p["img"]), stride).numpy(), p[key])


def test_exhaustive_voting():
    d = load("voting")
    grid = grids.Grid2D((12, 12), 0.2)
    t, tv = opv.sample_query_templates(d["fq"], d["vq"], 8, grid)
    assert np.array_equal(tv, d["t_valid"]); close(t, d["templates"], 1e-5)
    close(opv.template_matching(d["templates"], d["t_valid"], d["fm"], d["vm"]), d["scores"], 1e-5)
    close(opv.exhaustive_pose_voting(d["fq"], d["vq"], d["fm"], d["vm"], 8, grid, d["conf"]), d["scores_conf"], 1e-5)
    tf = opv.exhaustive_index_to_tfm(np.array([3, 14, 9]), grid, 8)
    close(tf.angle, d["tfm_angle"], 1e-6); close(tf.t, d["tfm_t"], 1e-5)
    close(opv.exhaustive_tfm_to_index(tf, grid, 8), d["index_back"], 1e-5)


@pytest.mark.parametrize("mode", ["max", "sum", "mean", "softmax", "weighted"])
def test_vertical_pooling_modes(mode):
    """oracle.bev_mapper.vertical_pooling vs the reference's own VerticalPooling.__call__ (tests/golden/make_golden_pooling.py)."""
    from oracle import bev_mapper as obm
    d = load("vertical_pooling")
    params = {"confidence_head": {"kernel": d["head_kernel"], "bias": d["head_bias"]}}
    out = obm.vertical_pooling(d["feats"], d["valid"], mode, params)
    plane, pvalid = out["plane"]
    assert np.array_equal(pvalid, d[f"{mode}_valid"]) and not pvalid[0] and not plane[0].any()
    close(plane, d[f"{mode}_plane"])
    if mode in ("softmax", "weighted"):
        close(out["scores"], d[f"{mode}_scores"])
        close(out["weights"], d[f"{mode}_weights"])
        assert not out["weights"][~d["valid"]].any()


@pytest.mark.parametrize("tag", ["allviews", "select"])
def test_lift_scene_vs_reference_streetview_encoder_call(tag):
    """oracle.bev_mapper.lift_scene (+ proj MLP) against the reference's OWN StreetViewEncoder.__call__
    (streetview_encoder.py:217-287) run under the stand-in (tests/golden/make_golden_sve.py): the whole lift orchestration,
    all-views path and view-selection path with max_view_distance."""
    from oracle import bev_mapper as obm
    d = load("sve_call_" + tag)
    tree = lambda pre: {n: {"kernel": d[f"{pre}_{n}_kernel"], "bias": d[f"{pre}_{n}_bias"]}
                        for n in ("Dense_0", "Dense_1") if f"{pre}_{n}_kernel" in d}
    f_proj = layers.mlp(d["f_img"][0], tree("proj"), apply_input_activation=True)          # :228-230
    close(f_proj[..., 8:], d["scores_images"][0])
    cam = geometry.Camera(wh=d["wh"][0], f=d["f"][0], c=d["c"][0]).scale(np.asarray([0.25, 0.25], F))   # :224
    T = geometry.Transform3D(R=d["R"][0], t=d["t"][0])
    mvd = float(d["max_view_distance"])
    f_grid, valid, _, _ = obm.lift_scene(f_proj, cam, T, d["xyz"][0], tree("fusion"), feature_dim=8, top_k=4,
                                         max_view_distance=None if mvd < 0 else mvd)
    assert np.array_equal(valid, d["valid"][0].astype(bool)) and 0.3 < valid.mean() < 0.95
    close(f_grid, d["volume"][0], tol=5e-5)
    assert not f_grid[~valid].any()


@pytest.mark.parametrize("tag", ["plain_allviews", "plain_select", "plain_depthmlp_allviews", "plain_depthmlp_select"])
def test_lift_scene_unweighted_vs_reference_streetview_encoder_call(tag):
    """The `do_weighted_fusion=False` branch (streetview_encoder.py:262-267: no proj MLP, plain mean / variance pooling,
    optional per-observation `depth_mlp` residual on [f, log10 depth, ray]) of oracle.bev_mapper.lift_scene against the
    reference's OWN StreetViewEncoder.__call__ run under the stand-in, all-views and view-selection paths."""
    from oracle import bev_mapper as obm
    d = load("sve_call_" + tag)
    tree = lambda pre: {n: {"kernel": d[f"{pre}_{n}_kernel"], "bias": d[f"{pre}_{n}_bias"]}
                        for n in ("Dense_0", "Dense_1") if f"{pre}_{n}_kernel" in d}
    assert d["scores_images"].size == 0 and d["fusion_Dense_0_kernel"].shape[0] == 16      # [mean | var], no score_max row
    cam = geometry.Camera(wh=d["wh"][0], f=d["f"][0], c=d["c"][0]).scale(np.asarray([0.25, 0.25], F))   # :224
    T = geometry.Transform3D(R=d["R"][0], t=d["t"][0])
    mvd = float(d["max_view_distance"])
    f_grid, valid, _, _ = obm.lift_scene(d["f_img"][0], cam, T, d["xyz"][0], tree("fusion"), feature_dim=8, top_k=4,
                                         max_view_distance=None if mvd < 0 else mvd, weighted=False,
                                         depth_mlp_params=tree("depth") if "depthmlp" in tag else None)
    assert np.array_equal(valid, d["valid"][0].astype(bool)) and 0.3 < valid.mean() < 0.95
    close(f_grid, d["volume"][0], tol=5e-5)
    assert not f_grid[~valid].any()


@pytest.mark.parametrize("tag", ["plain_allviews", "plain_select"])
def test_unweighted_lift_equals_weighted_lift_with_zero_logits(tag):
    """Design check of the product's un-weighted path (snap_b200/streetview_encoder.py): with all-zero scale logits the
    weighted soft-max is uniform over the valid views, so [mean | var] equal the plain statistics and the extra
    score_max column (= 0) meets a zero row appended to the fusion MLP's first kernel -- the weighted kernels then
    reproduce the reference's `do_weighted_fusion=False` volume."""
    from oracle import bev_mapper as obm
    d = load("sve_call_" + tag)
    fusion = {n: {"kernel": d[f"fusion_{n}_kernel"], "bias": d[f"fusion_{n}_bias"]} for n in ("Dense_0", "Dense_1")}
    fusion["Dense_0"]["kernel"] = np.concatenate([fusion["Dense_0"]["kernel"], np.zeros((1, 12), F)])
    f_pad = np.concatenate([d["f_img"][0], np.zeros(d["f_img"][0].shape[:-1] + (6,), F)], -1)
    cam = geometry.Camera(wh=d["wh"][0], f=d["f"][0], c=d["c"][0]).scale(np.asarray([0.25, 0.25], F))
    T = geometry.Transform3D(R=d["R"][0], t=d["t"][0])
    mvd = float(d["max_view_distance"])
    f_grid, valid, _, _ = obm.lift_scene(f_pad, cam, T, d["xyz"][0], fusion, feature_dim=8, top_k=4,
                                         max_view_distance=None if mvd < 0 else mvd)
    assert np.array_equal(valid, d["valid"][0].astype(bool))
    close(f_grid, d["volume"][0], tol=5e-5)


def test_bev_mapper_forward_vs_reference_bevmapper_call():
    """oracle.bev_mapper_forward (encoders bypassed) against the reference's OWN BEVMapper.__call__ chain
    (bev_mapper.py:159-296 + StreetViewEncoder.__call__ + VerticalPooling.__call__) run under the stand-in
    (tests/golden/make_golden_bevmapper.py): voxel grid from the median camera height, lift, vertical max, modality max
    with the aerial plane, matching head, confidence."""
    from oracle import bev_mapper as obm
    d = load("bevmapper_call")
    tree = lambda pre: {n: {"kernel": d[f"{pre}_{n}_kernel"], "bias": d[f"{pre}_{n}_bias"]}
                        for n in ("Dense_0", "Dense_1") if f"{pre}_{n}_kernel" in d}
    params = {"streetview_encoder": {"proj_mlp": tree("proj"), "fusion_mlp": tree("fusion")},
              "matching_proj": {"kernel": d["Wm"], "bias": d["bm"]}}
    data = {"camera": geometry.Camera(wh=d["wh"], f=d["f"], c=d["c"]), "T_view2scene": geometry.Transform3D(R=d["R"], t=d["t"])}
    grid = grids.Grid2D((6, 6), float(d["cell"]))
    pred = obm.bev_mapper_forward(data, params, grid, scene_z_offset=1.0, scene_z_height=1.6, feature_dim=8,
                                  precomputed={"sv_features": d["f_img"], "sv_stride": (4.0, 4.0), "aerial": d["aerial"]})
    for b in range(2):   # the voxel grid (bev_mapper.py:162-196): z offset = median camera height - scene_z_offset
        xyz, _ = obm.build_xyz_query(grid, d["t"][b], 1.0, 1.6)
        close(xyz, d["xyz"][b], tol=1e-6)
        sv = pred["streetview"][b]
        assert np.array_equal(sv["valid"], d["sv_valid"][b].astype(bool))
        close(sv["feature_plane"], d["sv_plane"][b], tol=5e-5)
    assert 0.3 < d["sv_valid"].mean() < 0.95
    assert np.array_equal(pred["bev_features"]["valid"], d["bev_valid"].astype(bool)) and d["bev_valid"].all()
    close(pred["bev_features"]["features"], d["bev_features"], tol=5e-5)
    close(pred["bev_matching"]["features"], d["bev_matching"], tol=5e-5)
    conf = obm.bev_confidence(pred["bev_features"]["features"], pred["bev_features"]["valid"],
                              {"layers_0": {"kernel": d["wc"], "bias": d["bc"]}})
    close(conf, d["bev_confidence"], tol=5e-5)


def test_image_encoder_wrapper_vs_reference_call(monkeypatch):
    """The wrapper logic of oracle.image_encoder.image_encoder (pad, level order, strides, crop) against the reference's OWN
    ImageEncoder.__call__ (image_encoder.py:119-144) run with a stand-in encoder / identity decoder
    (tests/golden/make_golden_image_encoder_call.py); the same stand-ins are patched into the oracle."""
    d = load("image_encoder_call")

    def pool(x, s):
        B, H, W, C = x.shape
        return x.reshape(B, H // s, s, W // s, s, C).mean((2, 4))
    for tag, skip_root in (("sv_30x44", False), ("sv_64x32", False), ("aerial_20x24", True)):
        base = 1 if skip_root else 4
        monkeypatch.setattr(ores, "resnet_v2", lambda img, p, skip, rd, trace=None, base=base: [pool(img, base * 2 ** k) for k in range(4)])
        monkeypatch.setattr(oie, "fpn_decoder", lambda feats, p, rd=None: feats)
        params = {"encoder": {f"block{k + 1}": {} for k in range(4)}, "decoder": {}}
        feats, strides = oie.image_encoder(torch.from_numpy(d[f"{tag}_image"]), params, skip_root)
        assert len(feats) == 4
        for k in range(4):
            ref = d[f"{tag}_feat{k}"]
            assert tuple(feats[k].shape) == ref.shape, (tag, k, feats[k].shape, ref.shape)
            close(feats[k].numpy(), ref, tol=1e-5)   # stand-in pooling: torch vs NumPy summation order
            assert tuple(float(x) for x in strides[k]) == tuple(float(x) for x in d[f"{tag}_stride{k}"])


def _golden_encoder_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_encoder", os.path.join(G, "make_golden_encoder.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # only defines the seeded case builders (the reference is not needed here)
    return mod


def _tt(tree):
    return {k: (_tt(v) if isinstance(v, dict) else torch.from_numpy(np.ascontiguousarray(v, dtype=F))) for k, v in tree.items()}


@pytest.mark.parametrize("tag", ["sv", "aerial"])
def test_image_encoder_vs_reference_modules(tag):
    """oracle.image_encoder (ResNetV2 + FPN, street-view and aerial variants) against the reference's OWN ResNetV2 /
    ResidualUnit / RootBlock / GroupNorm / StdConv / FPNDecoder / ImageEncoder modules executed under the jax + flax.linen
    stand-ins (tests/golden/make_golden_encoder.py); parameters are re-created from the fixture's seeds."""
    mod = _golden_encoder_module()
    d = load("encoder_modules")
    _, p, img, skip_root = mod.encoder_case(tag)
    feats, strides = oie.image_encoder(torch.from_numpy(img), _tt(p), skip_root)
    assert len(feats) == 4
    for k, f in enumerate(feats):
        ref = d[f"{tag}_feat{k}"]
        assert tuple(f.shape) == ref.shape
        close(f.numpy(), ref, tol=2e-4)      # fp32 conv summation order through ~20 layers with GroupNorm


def test_resnet_stage_and_mlp_vs_reference_modules():
    mod = _golden_encoder_module()
    d = load("encoder_modules")
    sp, x = mod.stage_case()
    y = ores.resnet_stage(torch.from_numpy(x), _tt(sp), 1)
    close(y.numpy(), d["stage_y"], tol=5e-5)
    mp, xm = mod.mlp_case()
    for act in (False, True):
        close(layers.mlp(xm, mp, apply_input_activation=act), d[f"mlp_y{int(act)}"], tol=1e-5)


@pytest.mark.parametrize("decoder_type", ["mlp", "resnet_stage"])
def test_semantic_head_vs_reference_semanticnet_call(decoder_type):
    """The semantic head (both decoder types) against the reference's OWN SemanticNet.__call__ (semantic_net.py:145-198:
    Dense -> ResNetStage -> MLP or the plain MLP, f32 logits zeroed where invalid, split into areas / exclusive / independent)
    executed under the stand-ins with a stand-in bev_mapper (tests/golden/make_golden_encoder.py)."""
    from oracle import semantic_net as osn
    mod = _golden_encoder_module()
    d = load("encoder_modules")
    cfg, p, feats, valid = mod.semantic_case(decoder_type)
    if decoder_type == "mlp":
        tp = {k: {n: torch.from_numpy(np.ascontiguousarray(v[n], dtype=F)) for n in v} for k, v in p.items()}
        logits = osn.mlp_head_forward_torch(torch.from_numpy(feats), valid, tp).numpy()
    else:
        logits = osn.semantic_decoder(feats, valid, p)
    assert logits.shape == (2, 6, 5, 12) and not logits[~valid].any()
    close(logits[..., :5], d[f"sem_{decoder_type}_areas"], tol=5e-5)
    close(logits[..., 5:9], d[f"sem_{decoder_type}_excl"], tol=5e-5)
    close(logits[..., 9:], d[f"sem_{decoder_type}_indep"], tol=5e-5)
