"""Golden fixtures for the semantic losses / metrics: the reference's OWN `snap/models/semantic_net.py:31-110`
(balancing_weights, multiclass_crossentropy_metrics, binary_crossentropy_metrics) and `layers.masked_mean`, executed under
the NumPy stand-in for jax / optax (tests/golden/jaxshim).  Run in the build container only:

    python tests/golden/make_golden_semantics.py     # writes tests/golden/sem_loss.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")

import jaxshim  # noqa: E402

jaxshim.install(REF)
from snap.models import semantic_net as sn  # noqa: E402

F = np.float32
rng = np.random.default_rng(77)
B, H, W = 3, 12, 10
area = ("crosswalk", "sidewalk", "road", "terrain", "building")
excl = ("fence", "pole", "tree", "void")
indep = ("traffic_sign", "traffic_light", "street_light")
la = rng.integers(0, len(area), (B, H, W))
le = rng.integers(0, len(excl), (B, H, W))
mi = rng.random((B, H, W, len(indep))) < 0.2
valid = rng.random((B, H, W)) < 0.7
valid[2] = False                                   # empty mask: masked_mean divides by the number of cells
logits_a = (rng.standard_normal((B, H, W, len(area))) * 2).astype(F)
logits_e = (rng.standard_normal((B, H, W, len(excl))) * 2).astype(F)
logits_i = (rng.standard_normal((B, H, W, len(indep))) * 2).astype(F)
fa = dict(zip(area, [0.02, 0.2, 0.5, 0.1, 0.3]))
fo = dict(zip(excl + indep, [0.01, 0.002, 0.05, 0.9, 0.0005, 0.0002, 0.003]))
d = dict(la=la, le=le, mi=mi, valid=valid, logits_a=logits_a, logits_e=logits_e, logits_i=logits_i,
         fa=np.array([fa[c] for c in area]), fo=np.array([fo[c] for c in excl + indep]))
for tag, (f1, f2) in {"plain": (None, None), "bal": (fa, fo)}.items():
    nll, m = sn.multiclass_crossentropy_metrics(logits_a, la, valid, area, f1)
    d[f"a_nll_{tag}"], d[f"a_acc_{tag}"], d[f"a_recall_{tag}"] = nll, m["accuracy"], np.stack([m[f"recall/{c}"] for c in area], -1)
    d[f"a_recall_avg_{tag}"] = m["recall/average"]
    nll, m = sn.multiclass_crossentropy_metrics(logits_e, le, valid, excl, f2, namespace="excl")
    d[f"e_nll_{tag}"], d[f"e_acc_{tag}"], d[f"e_recall_{tag}"] = nll, m["accuracy/excl"], np.stack([m[f"recall/{c}"] for c in excl], -1)
    nll, m = sn.binary_crossentropy_metrics(logits_i, mi, valid, indep, f2, namespace="indep")
    d[f"i_nll_{tag}"], d[f"i_recall_{tag}"] = nll, np.stack([m[f"recall/{c}"] for c in indep], -1)
    d[f"i_recall_avg_{tag}"] = m["recall/average/indep"]
# ---- label preparation (:254-298) through the reference's own methods on a stand-in `self` ---------------------------
import types as pytypes  # noqa: E402
gt_classes = ("road", "crosswalk", "sidewalk", "terrain", "building", "fence", "pole", "tree", "traffic_sign",
              "traffic_light", "street_light")
fake = pytypes.SimpleNamespace(
    gt_indices={c: i for i, c in enumerate(gt_classes)},
    config=pytypes.SimpleNamespace(area_classes=area, object_classes_exclusive=excl[:-1], object_classes_independent=indep))
fake._create_exclusive_labels = lambda *a, **k: sn.SemanticNetModel._create_exclusive_labels(fake, *a, **k)
gt_masks = rng.random((B, H, W, len(gt_classes))) < 0.25
lab_a, val_a = sn.SemanticNetModel.create_area_labels(fake, gt_masks)
lab_e, m_i = sn.SemanticNetModel.create_object_labels(fake, gt_masks)
d.update(gt_masks=gt_masks, lab_area=lab_a, valid_area=val_a, lab_excl=lab_e, masks_indep=m_i)
d["w_area"] = sn.balancing_weights(dict(fa), area)
wp, wn = sn.balancing_weights(dict(fo), indep, binary=True)
d["w_pos"], d["w_neg"] = wp, wn
np.savez_compressed(os.path.join(HERE, "sem_loss.npz"), **{k: np.asarray(v) for k, v in d.items()})
print({k: np.asarray(v).shape for k, v in d.items()})
