"""Known-answer tests of the CPU oracle — the oracle against independent implementations
(SciPy) and against analytic properties of the reference algorithm."""
import numpy as np
import pytest
import scipy.ndimage
import scipy.signal
import torch

from oracle import bev_mapper as obm, geometry, grids, image_encoder as oie, layers, pose_exhaustive_voting as opv
from oracle import resnet as ores, streetview_encoder as osv

F = np.float32


def test_interpolate_nd_equals_scipy_map_coordinates():
    """grids.py:125-130: interpolate_nd == map_coordinates(order=1, mode='nearest') at (p - 0.5)."""
    rng = np.random.default_rng(0)
    arr = rng.standard_normal((13, 17, 3)).astype(F)
    pts = (rng.random((500, 2)) * [15, 19] - 1).astype(F)
    val, valid = grids.interpolate_nd(arr, pts)
    for d in range(3):
        ref = scipy.ndimage.map_coordinates(arr[..., d], (pts - 0.5).T, order=1, mode="nearest")
        assert np.abs(val[:, d] - ref).max() < 1e-5
    assert np.array_equal(valid, ((pts >= 0) & (pts < [13, 17])).all(-1))


def test_nan_mask_validity_erodes_zero_weight_taps():
    """grids.py:131-136 / SURVEY A.3: a point exactly on a cell centre is invalid if a zero-weight tap is."""
    arr = np.ones((4, 4, 1), F)
    mask = np.ones((4, 4), bool)
    mask[2, 2] = False
    pts = np.array([[1.5, 1.5], [0.5, 0.5], [1.2, 1.2]], F)  # centre of (1,1): taps (1,1),(1,2),(2,1),(2,2)
    _, valid = grids.interpolate_nd(arr, pts, mask)
    assert valid.tolist() == [False, True, True]


def test_pinhole_projection_known_answers():
    """geometry.py:177,198-221: optical axis -> principal point; z < 1e-3 invisible; bounds half-open."""
    cam = geometry.Camera(wh=np.array([[160, 120]], F), f=np.array([[100, 100]], F), c=np.array([[80, 60]], F))
    p = np.array([[[0, 0, 5], [0, 0, 5e-4], [-4, 0, 5], [4, 0, 5], [0, 3.0, 5]]], F)
    p2d, vis = cam.world2image(p)
    assert np.allclose(p2d[0, 0], [80, 60]) and vis[0].tolist() == [True, False, True, False, False]
    assert p2d[0, 2, 0] == 0.0 and p2d[0, 3, 0] == 160.0 and p2d[0, 4, 1] == 120.0


def test_transform_inverse_roundtrip():
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    T = geometry.Transform3D(R=q.astype(F)[None], t=rng.standard_normal((1, 3)).astype(F))
    pts = rng.standard_normal((1, 10, 3)).astype(F)
    back = T.inv.transform(T.transform(pts))
    assert np.abs(back - pts).max() < 1e-5


def test_pad_to_multiple_pads_full_stride_when_divisible():
    """image_encoder.py:37 / SURVEY D1: 224->256, 480->512, 640->672."""
    for (h, w, s, eh, ew) in [(224, 224, 32, 256, 256), (480, 640, 32, 512, 672), (128, 128, 8, 136, 136)]:
        y = oie.pad_to_multiple(torch.ones((1, h, w, 3)), s)
        assert tuple(y.shape[1:3]) == (eh, ew) and float(y[0, h:, :, :].abs().sum()) == 0.0


def test_groupnorm_and_stdconv_known_answers():
    """resnet.py:34-79: GN of a constant input returns the bias; StdConv kernels have zero mean / unit RMS."""
    c = 64
    x = torch.full((2, 5, 7, c), 3.25)
    bias = torch.linspace(-1, 1, c)
    y = ores.group_norm(x, torch.ones(c) * 1.7, bias)
    assert torch.allclose(y, bias.expand_as(y), atol=1e-6)
    w = ores.std_kernel(torch.randn((3, 3, 16, 8)) * 5 + 2)
    assert w.mean(dim=[0, 1, 2]).abs().max() < 1e-5
    assert ((w * w).mean(dim=[0, 1, 2]) - 1).abs().max() < 1e-4


def test_depth_score_and_pooling_known_answers():
    """streetview_encoder.py:109-124,141-178."""
    S = 32
    scales = np.arange(S, dtype=F)[None]                   # logit of bin i is i
    d = np.array([0.5, 1.0, 32.0, 100.0, np.sqrt(32.0)], F)
    s = osv.interpolate_depth_score(np.repeat(scales, 5, 0), d)
    assert np.allclose(s, [0, 0, 31, 31, 15.5], atol=1e-4)
    feats = np.array([[[1.0, 2.0], [3.0, 6.0], [100.0, 100.0]]], F)   # N=1, V=3, D=2
    valid = np.array([[True, True, False]])
    scores = np.array([[0.0, 0.0, 50.0]], F)
    stats, any_ = osv.pool_multiview_features(feats, valid, scores, False, True)
    assert any_[0] and np.allclose(stats[0], [2.0, 4.0, 1.0, 4.0, 0.0])   # mean, var (w = 1/2), max score
    stats0, any0 = osv.pool_multiview_features(feats, np.zeros((1, 3), bool), scores, False, True)
    assert not any0[0] and not stats0.any()


def test_view_selection_ties_go_to_lower_index():
    """lax.top_k semantics, SURVEY A.5."""
    T = geometry.Transform3D(R=np.tile(np.eye(3, dtype=F), (4, 1, 1)), t=np.zeros((4, 3), F))
    vis = np.array([[True, False, True, True]])
    idx, md = osv.view_selection(np.ones((1, 3), F), T, vis, 3)
    assert idx[0].tolist() == [0, 2, 3] and np.isclose(md[0], np.sqrt(3))


def test_vertical_pooling_double_where():
    f = np.array([[[1.0, -5.0], [2.0, -7.0], [9.0, 9.0]], [[1, 1], [2, 2], [3, 3]]], F)  # 2 cells, Z=3, D=2
    v = np.array([[True, True, False], [False, False, False]])
    plane, pv = obm.vertical_pooling_max(f, v)
    assert pv.tolist() == [True, False] and plane.tolist() == [[2.0, -5.0], [0.0, 0.0]]


def test_vertical_pooling_modes_known_answers():
    """bev_mapper.py:56-88: sum / mean / softmax / weighted / mlp on hand-computable columns."""
    f = np.array([[[1.0, -5.0], [2.0, -7.0], [9.0, 9.0]], [[1, 1], [2, 2], [3, 3]]], F)  # 2 cells, Z=3, C=2
    v = np.array([[True, True, False], [False, False, False]])
    r = obm.vertical_pooling(f, v, "max")["plane"]
    assert r[1].tolist() == [True, False] and r[0].tolist() == [[2.0, -5.0], [0.0, 0.0]]
    assert obm.vertical_pooling(f, v, "sum")["plane"][0].tolist() == [[3.0, -12.0], [0.0, 0.0]]
    assert obm.vertical_pooling(f, v, "mean")["plane"][0].tolist() == [[1.5, -6.0], [0.0, 0.0]]
    # zero confidence kernel: every logit = bias; softmax -> uniform over the valid levels
    head = {"confidence_head": {"kernel": np.zeros((2, 1), F), "bias": np.array([0.7], F)}}
    r = obm.vertical_pooling(f, v, "softmax", head)
    assert np.allclose(r["scores"], 0.7) and np.allclose(r["weights"], [[0.5, 0.5, 0], [0, 0, 0]])
    assert np.allclose(r["plane"][0], [[1.5, -6.0], [0, 0]])
    # 'weighted': logits go through log_sigmoid; kernel picks channel 0 -> logits (1, 2, 9)
    head = {"confidence_head": {"kernel": np.array([[1.0], [0.0]], F), "bias": np.zeros(1, F)}}
    r = obm.vertical_pooling(f, v, "weighted", head)
    ls = -np.log1p(np.exp(-np.array([1.0, 2.0])))
    w = np.exp(ls - 0.0) / np.exp(ls - 0.0).sum()     # shift = max(0, max ls) = 0 since log_sigmoid < 0
    assert np.allclose(r["scores"][0, :2], ls, atol=1e-6) and np.allclose(r["weights"][0], [w[0], w[1], 0], atol=1e-6)
    assert np.allclose(r["plane"][0][0], [w[0] * 1 + w[1] * 2, w[0] * -5 + w[1] * -7], atol=1e-5)
    # 'mlp': the invalid level is zeroed before the flatten, so a sum-all kernel gives the sum over valid levels
    mlp = {"fusion_mlp": {"Dense_0": {"kernel": np.ones((6, 1), F), "bias": np.array([0.5], F)}}}
    r = obm.vertical_pooling(f, v, "mlp", mlp)
    assert r["plane"][0].tolist() == [[1 - 5 + 2 - 7 + 0.5], [0.0]]
    conf = obm.bev_confidence(np.array([[1.0, 0.0], [3.0, 0.0]], F), np.array([True, False]),
                              {"layers_0": {"kernel": np.array([[1.0], [0.0]], F), "bias": np.zeros(1, F)}})
    assert np.allclose(conf, [-np.log1p(np.exp(-1.0)), 0.0], atol=1e-6)


def test_normalize_zero_vector():
    x = np.array([[3.0, 4.0], [0.0, 0.0], [1e-7, 0.0]], F)
    assert np.allclose(layers.normalize(x), [[0.6, 0.8], [0, 0], [0, 0]])


@pytest.mark.parametrize("G,R", [(16, 8), (24, 12)])
def test_templates_quadrant_and_identity_properties(G, R):
    """pose_exhaustive_voting.py:56-68."""
    rng = np.random.default_rng(2)
    g = grids.Grid2D((G, G), 0.2)
    f = rng.standard_normal((G, G, 5)).astype(F)
    v = np.ones((G, G), bool)
    v[:3, :4] = False
    t, tv = opv.sample_query_templates(f, v, R, g)
    nq = R // 4
    for k in range(1, 4):
        assert np.array_equal(t[k * nq:(k + 1) * nq], np.rot90(t[:nq], k, axes=(2, 1)))
        assert np.array_equal(tv[k * nq:(k + 1) * nq], np.rot90(tv[:nq], k, axes=(2, 1)))
    assert np.abs(np.where(tv[0][..., None], t[0] - f, 0)).max() < 1e-4      # rotation 0 = identity where valid
    assert not (tv[0] & ~v).any()


def test_template_matching_equals_scipy_and_peaks_at_identity():
    """pose_exhaustive_voting.py:83-103 vs scipy.signal.convolve, channel by channel."""
    rng = np.random.default_rng(3)
    G, R, D = 10, 4, 3
    g = grids.Grid2D((G, G), 0.2)
    m = rng.standard_normal((G, G, D)).astype(F)
    mv = rng.random((G, G)) > 0.1
    q, qv = opv.sample_query_templates(m, np.ones((G, G), bool), R, g)
    got = opv.template_matching(q, qv, m, mv)
    m_pad = np.pad(m, ((G - 1, G - 1), (G - 1, G - 1), (0, 0)), mode="edge")
    mv_pad = np.pad(mv, ((G - 1, G - 1), (G - 1, G - 1)), mode="constant")
    for r in range(R):
        sc = sum(scipy.signal.convolve(q[r, ::-1, ::-1, d], m_pad[..., d], mode="valid") for d in range(D))
        nv = scipy.signal.convolve(qv[r].astype(F), mv_pad.astype(F), mode="valid", method="direct")  # exact integers
        ref = np.where(nv > F(0.05 * G * G), sc, -np.inf) / qv[r].sum()
        fin = np.isfinite(ref)
        assert np.array_equal(fin, np.isfinite(got[r]))
        assert np.abs(got[r][fin] - ref[fin]).max() < 1e-4
    s = opv.template_matching(q, qv, m, np.ones((G, G), bool))
    assert np.unravel_index(np.argmax(np.where(np.isfinite(s), s, -np.inf)), s.shape) == (0, G - 1, G - 1)


def test_pose_index_transform_are_inverse():
    g = grids.Grid2D((64, 64), 0.2)
    for idx in ([0, 63, 63], [5, 10, 100], [35, 126, 0]):
        tf = opv.exhaustive_index_to_tfm(np.array(idx), g, 36)
        back = opv.exhaustive_tfm_to_index(tf, g, 36)
        assert np.allclose(back, idx, atol=2e-3)
