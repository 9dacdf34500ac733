"""Host-side logic of the product package (no GPU): geometry/grid preparation must be BIT-identical to the
oracle (it feeds the bit-exact visibility kernels), configs mirror defaults.py, parameter trees carry the
Flax names of SURVEY.md Appendix B."""
import numpy as np
import torch

from oracle import bev_mapper as obm, geometry as ogeo, grids as ogrids, pose_exhaustive_voting as opv
from snap_b200 import _lib, bev_mapper, configs, params, pose_exhaustive_voting as pv, streetview_encoder as sve
from snap_b200 import synthetic, types

F = np.float32


def test_xyz_grid_bit_identical_to_oracle():
    for G, V in [(64, 1), (128, 4), (32, 3)]:
        data = synthetic.make_tile(3, V, (96, 128), G, batch=2)
        mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
        xs, ys, zs = mapper.build_xyz_grid(data)
        for b in range(2):
            xyz, zoff = obm.build_xyz_query(ogrids.Grid2D((G, G), 0.2), data["T_view2scene"].t[b])
            assert np.array_equal(xyz[:, 0, 0, 0], xs) and np.array_equal(xyz[0, :, 0, 1], ys)
            assert np.array_equal(xyz[0, 0, :, 2], zs[b]) and zs.shape[1] == 60


def test_view_table_bit_identical_to_oracle_inverse_and_scale():
    data = synthetic.make_tile(5, 4, (480, 640), 128, fisheye=True)
    words = sve.pack_views(data["camera"], data["T_view2scene"], 0, (4.0, 4.0))
    raw = words.tobytes()
    T = ogeo.Transform3D(R=data["T_view2scene"].R[0], t=data["T_view2scene"].t[0]).inv
    cam = ogeo.Camera(wh=data["camera"].wh[0], f=data["camera"].f[0], c=data["camera"].c[0]).scale(np.asarray([0.25, 0.25], F))
    for v in range(4):
        lv = _lib.LiftView.from_buffer_copy(raw[v * 92:(v + 1) * 92])
        assert np.array_equal(np.array(lv.Rinv[:], F).reshape(3, 3), T.R[v])
        assert np.array_equal(np.array(lv.tinv[:], F), T.t[v])
        assert np.array_equal(np.array(lv.f[:], F), cam.f[v]) and np.array_equal(np.array(lv.wh[:], F), cam.wh[v])
        assert lv.fisheye == 1 and np.isclose(lv.tan_half_fov, np.tan(np.deg2rad(57.5)), rtol=1e-6)


def test_template_rotation_params_bit_identical_to_oracle():
    for G, R in [(128, 36), (64, 8)]:
        rot = pv.template_rotation_params(R, types.Grid2D((G, G), 0.2))
        t = opv.template_transforms(R, ogrids.Grid2D((G, G), 0.2))
        nq = R // 4
        ref = np.stack([np.cos(t.angle[:nq]), np.sin(t.angle[:nq]), t.t[:nq, 0], t.t[:nq, 1]], -1).astype(F)
        assert np.array_equal(rot, ref)


def test_rot90_index_map_of_the_kernel():
    """csrc/xcorr.cu rot90_src: template quadrant k at (a,b) reads the quarter at (i,j)."""
    G = 7
    base = np.arange(G * G).reshape(G, G)

    def src(k, a, b):
        return [(a, b), (G - 1 - b, a), (G - 1 - a, G - 1 - b), (b, G - 1 - a)][k]
    for k in range(4):
        ref = np.rot90(base[None], k, axes=(2, 1))[0]
        got = np.array([[base[src(k, a, b)] for b in range(G)] for a in range(G)])
        assert np.array_equal(got, ref), k


def test_pose_index_roundtrip_product_side():
    g = types.Grid2D((128, 128), 0.2)
    idx = pv.exhaustive_tfm_to_index(*pv.exhaustive_index_to_tfm(np.array([7, 200, 31]), g, 36), g, 36)
    assert np.allclose(idx, [7, 200, 31], atol=2e-3)


def test_configs_mirror_reference_defaults():
    c = configs.bev_mapper()
    sv = c.streetview_encoder
    assert (sv.feature_dim, sv.num_scale_bins, sv.top_k_view_selection, tuple(sv.depth_min_max)) == (128, 32, 4, (1.0, 32.0))
    assert tuple(sv.fusion.layers) == (256, 128) and sv.proj_mlp.apply_input_activation and sv.do_weighted_fusion
    assert (c.scene_z_offset, c.scene_z_height, c.matching_dim, c.pooling.pooling) == (4.0, 12.0, 32, "max")
    assert c.aerial_encoder.encoder.skip_root_block and not sv.image_encoder.encoder.skip_root_block
    assert configs.get_block_desc(50) == [3, 4, 6, 3]


def test_param_tree_names_and_shapes():
    p = params.init_bev_mapper(np.random.default_rng(0), configs.bev_mapper())
    enc = p["streetview_encoder"]["image_encoder"]["encoder"]
    assert enc["root_block"]["conv_root"]["kernel"].shape == (7, 7, 3, 64)
    assert enc["block1"]["unit01"]["conv_proj"]["kernel"].shape == (1, 1, 64, 256)
    assert enc["block4"]["unit03"]["conv2"]["kernel"].shape == (3, 3, 512, 512)
    assert "conv_proj" not in enc["block2"]["unit02"] and enc["block3"]["unit06"]["gn1"]["scale"].shape == (1, 1, 1, 1024)
    dec = p["streetview_encoder"]["image_encoder"]["decoder"]
    assert dec["0_skip_conv"]["kernel"].shape == (1, 1, 2048, 128) and dec["3_skip_norm"]["bias"].shape == (1, 1, 1, 256)
    assert p["streetview_encoder"]["proj_mlp"]["Dense_0"]["kernel"].shape == (128, 160)
    assert p["streetview_encoder"]["fusion_mlp"]["Dense_0"]["kernel"].shape == (257, 256)
    assert p["aerial_encoder"]["encoder"]["conv_root"]["kernel"].shape == (3, 3, 3, 64)
    assert p["matching_proj"]["kernel"].shape == (128, 32)
    n = sum(x.size for x in _leaves(p))
    assert 47.5e6 < n < 48.5e6


def _leaves(t):
    for v in t.values():
        if isinstance(v, dict):
            yield from _leaves(v)
        else:
            yield v


def test_unsupported_configs_fail_loudly():
    import pytest
    c = configs.bev_mapper()
    c.streetview_encoder.do_weighted_fusion = False
    c.streetview_encoder.depth_mlp = configs.mlp()       # the per-observation depth_mlp residual (:263-267) needs its layers
    with pytest.raises(NotImplementedError):
        bev_mapper.BEVMapper(c, types.Grid2D((8, 8), 0.2))
    c.streetview_encoder.depth_mlp.layers = (64, 128)
    assert bev_mapper.BEVMapper(c, types.Grid2D((8, 8), 0.2)).streetview_encoder.has_depth_mlp
    c.streetview_encoder.do_weighted_fusion = True      # the weighted branch never builds a depth_mlp (:207-215)
    assert not bev_mapper.BEVMapper(c, types.Grid2D((8, 8), 0.2)).streetview_encoder.has_depth_mlp
    c3 = configs.bev_mapper()
    c3.streetview_encoder.fusion_add_minmax = True       # built on the unfused lift: [mean | var | max | min | score_max]
    m = bev_mapper.BEVMapper(c3, types.Grid2D((8, 8), 0.2))
    assert m.streetview_encoder.stats_dim == 513 and m.streetview_encoder.stats_ld == 544
    assert not m.streetview_encoder.default_stats
    c2 = configs.image_encoder()
    c2.encoder_name = "vit"
    with pytest.raises(ValueError):
        from snap_b200 import image_encoder
        image_encoder.ImageEncoder(c2)


def test_xy_bev_points_separable_and_paired():
    """data['xy_bev'] (bev_mapper.py:162-166): a regular (separable) point grid keeps the xs[X], ys[Y] layout, arbitrary
    points (the field-of-view filtered query frustum [N,1,2] of BEVLocalizer) switch the lift to per-column coordinates."""
    import pytest
    from snap_b200 import bev_localizer
    G = 16
    data = synthetic.make_tile(3, 1, (96, 128), G, batch=2)
    mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
    # separable: a shifted regular grid
    xs0, ys0 = types.Grid2D((6, 4), 0.5).cell_centers(0) - F(1.5), types.Grid2D((6, 4), 0.5).cell_centers(1)
    xy = np.stack(np.meshgrid(xs0, ys0, indexing="ij"), -1).astype(F)
    d = dict(data, xy_bev=xy)
    xs, ys, zs = mapper.build_xyz_grid(d)
    assert "xy_shape" not in d and np.array_equal(xs, xs0) and np.array_equal(ys, ys0) and zs.shape == (2, 60)
    # batched copies of the same grid are one launch; different grids per example are split by BEVMapper.apply into
    # single-example launches (tests/test_localizer_gpu.py) and never reach build_xyz_grid as a batch
    d = dict(data, xy_bev=np.stack([xy, xy]))
    assert np.array_equal(mapper.build_xyz_grid(d)[0], xs0)
    with pytest.raises(ValueError):
        mapper.build_xyz_grid(dict(data, xy_bev=np.stack([xy, xy + F(0.1)])))
    # arbitrary points: the query frustum of the localizer
    _, _, q = bev_localizer.build_query_frustum_grid(0.2, 16.0, True, 72.0)
    d = dict(data, xy_bev=q)
    xs, ys, _ = mapper.build_xyz_grid(d)
    assert d["xy_shape"] == (4652, 1) and np.array_equal(xs, q[:, 0, 0]) and np.array_equal(ys, q[:, 0, 1])
    lp = sve.fill_lift_params(configs.streetview_encoder(), 1, 24, 32, 4652, 1, 60, 288, True)
    assert lp.xy_paired == 1 and lp.no_variance == 0 and lp.add_minmax == 0 and lp.X == 4652 and lp.Y == 1


def test_localizer_pose_helpers():
    """Transform3D -> Transform2D of the ground truth (geometry.py:103-111) and the refinement lattice axes (pose_estimation.py:177-184)."""
    from oracle import pose_estimation as ope
    from snap_b200 import bev_localizer, pose_estimation
    a = np.array([0.3, -2.0, 3.1], F)
    R = np.stack([[[np.cos(x), -np.sin(x), 0], [np.sin(x), np.cos(x), 0], [0, 0, 1]] for x in a]).astype(F)
    t = np.arange(9, dtype=F).reshape(3, 3)
    out = bev_localizer.transform2d_from_transform3d(types.Transform3D(R=R, t=t))
    ang, tt = ope.transform2d_from_transform3d(R, t)
    assert np.array_equal(out[:, 0], ang) and np.array_equal(out[:, 1:], tt) and np.allclose(out[:, 0], a, atol=1e-6)
    rot, pos = pose_estimation.refinement_offsets()
    shape, off = ope.refinement_offsets()
    assert (len(rot), len(pos), len(pos)) == tuple(shape) == (41, 41, 41)
    assert np.array_equal(off[:, 0].reshape(shape)[:, 0, 0], rot) and np.array_equal(off[:, 2].reshape(shape)[0, 0, :], pos)
    assert rot[20] == 0 and abs(pos[20]) < 1e-6


def test_recover_dense_feature_plane():
    """bev_localizer.py:111-129: the field-of-view points scatter back onto the dense 120 x 80 frustum grid."""
    from snap_b200 import bev_localizer
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov = True
    cfg.num_pose_samples = 16
    loc = bev_localizer.BEVLocalizer(cfg, None, types.Grid2D((32, 32), 0.2))
    N = loc.q_xy_p.shape[0]
    feats = torch.arange(N * 3, dtype=torch.float32).reshape(N, 1, 3) + 1
    valid = torch.ones((N, 1), dtype=torch.uint8)
    dense = loc.recover_dense_feature_plane(types.FeaturePlane(feats, valid))
    assert tuple(dense.features.shape) == (120, 80, 3) and int(dense.valid.sum()) == N == 4652
    # point n sits in the cell its coordinates fall into; cells outside the field of view stay empty
    q = loc.q_xy_p[:, 0] + loc.qgrid_p_q
    ij = np.floor(q / F(0.2)).astype(int)
    assert np.array_equal(dense.features[ij[:, 0], ij[:, 1]].numpy(), feats[:, 0].numpy())
    assert not dense.features[0, 79].any() and not dense.valid[0, 79]        # far corner: |angle| > 36 degrees


def test_explicit_xyz_query_is_split_into_the_kernel_layout():
    """data['xyz_query'] (bev_mapper.py:162): the tensor the reference builds (:187-196) decomposes exactly into the
    (xs, ys, zs) the kernels take; a tensor without column structure is refused."""
    import pytest
    G = 8
    data = synthetic.make_tile(3, 2, (96, 128), G, batch=2)
    mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
    xs, ys, zs = mapper.build_xyz_grid(dict(data))
    xyz = np.zeros((2, G, G, zs.shape[1], 3), F)
    xyz[..., 0] = xs[None, :, None, None]
    xyz[..., 1] = ys[None, None, :, None]
    xyz[..., 2] = zs[:, None, None, :]
    d = dict(data, xyz_query=xyz)
    xs2, ys2, zs2 = mapper.split_xyz_query(d)
    assert np.array_equal(xs2, xs) and np.array_equal(ys2, ys) and np.array_equal(zs2, zs) and "xy_shape" not in d
    # arbitrary (x, y) per column: paired layout
    xyz2 = xyz.copy()
    xyz2[:, 3, 4, :, 0] += F(0.05)
    d = dict(data, xyz_query=xyz2)
    xs3, ys3, _ = mapper.split_xyz_query(d)
    assert d["xy_shape"] == (G, G) and xs3.shape == (G * G,) and xs3[3 * G + 4] == xyz2[0, 3, 4, 0, 0]
    bad = xyz.copy()
    bad[0, 1, 1, 5, 2] += F(0.1)          # one column with its own z level
    with pytest.raises(NotImplementedError):
        mapper.split_xyz_query(dict(data, xyz_query=bad))


def test_unweighted_fusion_configuration_host_side():
    """do_weighted_fusion=False (streetview_encoder.py:196-215): the module has no proj_mlp, the fusion MLP's first kernel
    has no score_max row; the product keeps the kernels' 257-wide statistics rows and refuses the depth_mlp residual."""
    import pytest
    from snap_b200 import configs, params, streetview_encoder as sve
    cfg = configs.streetview_encoder()
    cfg.do_weighted_fusion = False
    tree = params.init_streetview_encoder(np.random.default_rng(0), cfg)
    assert "proj_mlp" not in tree and tree["fusion_mlp"]["Dense_0"]["kernel"].shape == (256, 256)
    enc = sve.StreetViewEncoder(cfg)
    assert not enc.weighted and enc.stats_dim == 257 and enc.stats_ld == 288 and enc.default_stats
    cfg.fusion_add_minmax = True
    assert params.init_streetview_encoder(np.random.default_rng(0), cfg)["fusion_mlp"]["Dense_0"]["kernel"].shape == (512, 256)
    cfg.depth_mlp = configs.mlp()
    with pytest.raises(NotImplementedError):      # no layer widths given
        sve.StreetViewEncoder(cfg)
    cfg.depth_mlp.layers = (64, 128)
    assert sve.StreetViewEncoder(cfg).has_depth_mlp
    cfg.depth_mlp = None
    # the default (weighted) tree is unchanged: [mean | var | score_max] rows and a 128 -> 160 proj MLP
    tree = params.init_streetview_encoder(np.random.default_rng(0), configs.streetview_encoder())
    assert tree["proj_mlp"]["Dense_0"]["kernel"].shape == (128, 160) and tree["fusion_mlp"]["Dense_0"]["kernel"].shape == (257, 256)


def test_localizer_trainer_host_side_layout_and_clipping():
    """`localizer_trainer.LocalizerTrainer` without a GPU: leaf order / master layout, the clip factor of
    `jax.example_libraries.optimizers.clip_grads`, refusal of configurations whose gradient would be incomplete."""
    import numpy as np
    import pytest
    import torch
    from snap_b200 import bev_localizer, configs, localizer_trainer, params, types
    assert localizer_trainer.clip_scale(0.3, 1.0) == 1.0 and abs(localizer_trainer.clip_scale(4.0, 1.0) - 0.25) < 1e-12
    rng = np.random.default_rng(0)
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov = True
    cfg.num_pose_samples = 10
    loc = bev_localizer.BEVLocalizer(cfg, None, types.Grid2D((32, 32), 0.2))
    sv = {"proj_mlp": params.init_mlp(rng, 128, (160,)), "fusion_mlp": params.init_mlp(rng, 257, (256, 128)),
          "image_encoder": {"frozen": True}}
    p = loc.init_params({"streetview_encoder": sv, "matching_proj": {"kernel": np.ones((128, 32), np.float32),
                                                                     "bias": np.zeros(32, np.float32)}})
    tr = localizer_trainer.LocalizerTrainer(loc, p, device="cpu")
    assert tr.paths[-1] == ("temperature",) and len(tr.paths) == 9
    n = sum(int(np.prod(np.asarray(tr._get(p, q)).shape)) for q in tr.paths)
    assert n == 128 * 160 + 160 + 257 * 256 + 256 + 256 * 128 + 128 + 128 * 32 + 32 + 1
    assert tr.masters.flat.numel() >= n and float(tr.masters.views[-1][0]) == 2.0
    tr.masters.views[6].mul_(3.0)          # matching_proj kernel
    tr._rebuild_params()
    assert tr.params["bev_mapper"]["streetview_encoder"]["image_encoder"] is sv["image_encoder"]
    assert tr.params["bev_mapper"]["matching_proj"]["kernel"][0, 0] == 3.0 and p["bev_mapper"]["matching_proj"]["kernel"][0, 0] == 1.0
    assert tr.params["bev_mapper"]["streetview_encoder"]["fusion_mlp"]["Dense_0"]["kernel"].shape == (257, 256)
    cfg2 = configs.bev_localizer()
    cfg2.bev_mapper = configs.bev_mapper(("streetview",))
    cfg2.filter_points_in_fov = True
    cfg2.add_confidence_query = True
    with pytest.raises(NotImplementedError):
        localizer_trainer.LocalizerTrainer(bev_localizer.BEVLocalizer(cfg2, None, types.Grid2D((32, 32), 0.2)), p, device="cpu")


def test_batch_mask_rescales_the_loss_gradient_rows():
    """`trainer.py:221` means the loss over batch['batch_mask']: example b's cotangent rows get mask_b * B / sum(mask)."""
    import numpy as np
    import pytest
    import torch
    from snap_b200.semantic_net import apply_batch_mask
    d = torch.arange(24, dtype=torch.float32).reshape(12, 2).to(torch.bfloat16)
    ref = d.clone()
    apply_batch_mask(d, {}, 3)
    assert torch.equal(d, ref)
    apply_batch_mask(d, {"batch_mask": np.array([1, 1, 1])}, 3)
    assert torch.equal(d, ref)
    apply_batch_mask(d, {"batch_mask": np.array([1, 0, 1])}, 3)
    want = ref.float().view(3, 4, 2) * torch.tensor([1.5, 0.0, 1.5]).view(3, 1, 1)
    assert torch.equal(d.float().view(3, 4, 2), want.to(torch.bfloat16).float())
    with pytest.raises(ValueError):
        apply_batch_mask(d, {"batch_mask": np.array([0, 0, 0])}, 3)
