"""Semantic losses / metrics of the oracle (oracle/semantic_net.py) against the fixture produced by the REFERENCE'S OWN
snap/models/semantic_net.py:31-110 under the NumPy stand-in for jax / optax (tests/golden/make_golden_semantics.py)."""
import os

import numpy as np

from oracle import semantic_net as osn

F = np.float32
AREA = ("crosswalk", "sidewalk", "road", "terrain", "building")
EXCL = ("fence", "pole", "tree", "void")
INDEP = ("traffic_sign", "traffic_light", "street_light")


def load():
    return dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sem_loss.npz")))


def close(a, b, tol=2e-5):
    assert np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() <= tol * (1 + np.abs(b).max())


def test_losses_and_metrics_match_the_reference_source():
    d = load()
    fa = dict(zip(AREA, d["fa"]))
    fo = dict(zip(EXCL + INDEP, d["fo"]))
    close(osn.balancing_weights(fa, AREA), d["w_area"], 1e-6)
    wp, wn = osn.balancing_weights(fo, INDEP, binary=True)
    close(wp, d["w_pos"], 1e-6)
    close(wn, d["w_neg"], 1e-6)
    for tag in ("plain", "bal"):
        wa = osn.balancing_weights(fa, AREA) if tag == "bal" else None
        we = osn.balancing_weights(fo, EXCL) if tag == "bal" else None
        wpn = osn.balancing_weights(fo, INDEP, binary=True) if tag == "bal" else (None, None)
        nll, acc, rec = osn.multiclass_crossentropy_metrics(d["logits_a"], d["la"], d["valid"], wa)
        close(nll, d[f"a_nll_{tag}"]); close(acc, d[f"a_acc_{tag}"]); close(rec, d[f"a_recall_{tag}"])
        close(rec.mean(-1), d[f"a_recall_avg_{tag}"])
        nll, acc, rec = osn.multiclass_crossentropy_metrics(d["logits_e"], d["le"], d["valid"], we)
        close(nll, d[f"e_nll_{tag}"]); close(acc, d[f"e_acc_{tag}"]); close(rec, d[f"e_recall_{tag}"])
        nll, rec = osn.binary_crossentropy_metrics(d["logits_i"], d["mi"], d["valid"], *wpn)
        close(nll, d[f"i_nll_{tag}"]); close(rec, d[f"i_recall_{tag}"]); close(rec.mean(-1), d[f"i_recall_avg_{tag}"])
    assert d["a_nll_plain"][2] == 0 and d["i_nll_plain"][2] == 0      # empty mask -> zero (masked_mean)


def test_exclusive_labels():
    gt = ("road", "sidewalk", "line", "stopline", "otherlanemarking", "tree")
    m = np.zeros((2, 3, len(gt)), bool)
    m[0, 0, 0] = m[0, 0, 1] = True      # two labels: the first selected class wins (argmax)
    m[0, 1, 3] = True                   # stopline counts as 'line'
    m[1, 2, 5] = True
    labels, valid = osn.create_exclusive_labels(m, gt, ("sidewalk", "road", "line"))
    assert labels[0, 0] == 0 and valid[0, 0] and labels[0, 1] == 2 and valid[0, 1] and not valid[1, 2]
    labels, valid = osn.create_exclusive_labels(m, gt, ("tree",), add_void=True)
    assert labels[1, 2] == 0 and labels[0, 0] == 1 and not valid[0, 0]


def test_torch_loss_restatement_equals_numpy_oracle():
    """oracle.semantic_net.total_loss_torch (the differentiable form used as gradient oracle) == loss_metrics (pinned above)."""
    import torch
    d = load()
    fa = dict(zip(AREA, d["fa"]))
    fo = dict(zip(EXCL + INDEP, d["fo"]))
    gt_classes = AREA + EXCL[:-1] + INDEP
    rng = np.random.default_rng(5)
    B, H, W = d["la"].shape
    masks = rng.random((B, H, W, len(gt_classes))) < 0.3
    bev_valid = d["valid"]
    la, va = osn.create_exclusive_labels(masks, gt_classes, AREA)
    le, _ = osn.create_exclusive_labels(masks, gt_classes, EXCL[:-1], add_void=True)
    gi = {c: i for i, c in enumerate(gt_classes)}
    mi = masks[..., [gi[c] for c in INDEP]]
    logits = np.concatenate([d["logits_a"], d["logits_e"], d["logits_i"]], -1)
    for bal in (False, True):
        ol, _ = osn.loss_metrics(d["logits_a"], d["logits_e"], d["logits_i"], bev_valid, masks, gt_classes, AREA, EXCL[:-1],
                                 INDEP, fa if bal else None, fo if bal else None)
        w = (osn.balancing_weights(fa, AREA), osn.balancing_weights(fo, EXCL), *osn.balancing_weights(fo, INDEP, binary=True)) \
            if bal else (None,) * 4
        mean, total = osn.total_loss_torch(torch.from_numpy(logits), la, va, le, mi, bev_valid, len(AREA), len(EXCL), *w)
        close(total.numpy(), ol["total"])
        assert abs(float(mean) - float(ol["total"].mean())) < 1e-5


def test_label_preparation_matches_the_reference_methods():
    """create_area_labels / create_object_labels (semantic_net.py:254-298) of the reference, run on a stand-in `self`."""
    d = load()
    gt = AREA[2:3] + AREA[0:1] + AREA[1:2] + AREA[3:] + EXCL[:-1] + INDEP    # road, crosswalk, sidewalk, terrain, building, ...
    la, va = osn.create_exclusive_labels(d["gt_masks"], gt, AREA)
    assert np.array_equal(la, d["lab_area"]) and np.array_equal(va, d["valid_area"])
    le, _ = osn.create_exclusive_labels(d["gt_masks"], gt, EXCL[:-1], add_void=True)
    assert np.array_equal(le, d["lab_excl"]) and (le == 3).any()
    gi = {c: i for i, c in enumerate(gt)}
    assert np.array_equal(d["gt_masks"][..., [gi[c] for c in INDEP]], d["masks_indep"])
