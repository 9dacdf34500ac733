"""Semantic head forward vs oracle/semantic_net.py (bf16-emulation mode)."""
import numpy as np
import pytest
import torch

from util import F, bf16_np, rd_bf16, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_semantic_head_vs_oracle():
    from oracle import semantic_net as osn
    from snap_b200 import configs, params, semantic_net, types
    rng = np.random.default_rng(5)
    cfg = configs.semantic_net()
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_semantic_decoder(rng, cfg)))
    B, G = 2, 24
    feats = bf16_np(rng.standard_normal((B, G, G, 128)))
    valid = rng.random((B, G, G)) > 0.2
    feats *= valid[..., None]
    ref = osn.semantic_decoder(feats, valid, p, rd_bf16)
    ref32 = osn.semantic_decoder(feats, valid, p)
    head = semantic_net.SemanticHead(cfg)
    plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16).cuda(), torch.from_numpy(valid.astype(np.uint8)).cuda())
    pred = head.apply({"params": {"decoder": p}}, plane)
    torch.cuda.synchronize()
    got = torch.cat([pred["logits_areas"], pred["logits_objects_exclusive"], pred["logits_objects_independent"]], -1).cpu().numpy()
    assert got.shape == ref.shape == (B, G, G, 12) and got.dtype == np.float32
    assert pred["logits_areas"].shape[-1] == 5 and pred["logits_objects_exclusive"].shape[-1] == 4
    assert not got[~valid].any()
    e, e32, eref = rel_l2(got, ref), rel_l2(got, ref32), rel_l2(ref, ref32)
    print(f"semantic head: vs bf16-oracle {e:.4f}, vs fp32-oracle {e32:.4f}, bf16-oracle vs fp32-oracle {eref:.4f}")
    # two residual units + 3 dense layers with identical rounding points: bf16 flips only
    assert e < 1e-2


@pytest.mark.parametrize("balanced", [False, True])
def test_semantic_losses_against_oracle(balanced):
    """loss_metrics_function (semantic_net.py:300-343) on the GPU vs the oracle (pinned by tests/test_golden_semantics.py)."""
    from oracle import semantic_net as osn
    from snap_b200 import configs, semantic_net, types
    rng = np.random.default_rng(12)
    B, G = 3, 64
    cfg = configs.semantic_net()
    gt_classes = ("road", "crosswalk", "sidewalk", "terrain", "building", "fence", "pole", "tree", "traffic_sign",
                  "traffic_light", "street_light", "line")
    if balanced:
        cfg.area_frequencies = tuple(zip(cfg.area_classes, (0.02, 0.2, 0.5, 0.1, 0.3)))
        cfg.object_frequencies = tuple(zip((*cfg.object_classes_exclusive, "void", *cfg.object_classes_independent),
                                           (0.01, 0.002, 0.05, 0.9, 0.0005, 0.0002, 0.003)))
    masks = rng.random((B, G, G, len(gt_classes))) < 0.25
    bev_valid = rng.random((B, G, G)) < 0.7
    bev_valid[2] = False
    la = (rng.standard_normal((B, G, G, 5)) * 2).astype(F)
    le = (rng.standard_normal((B, G, G, 4)) * 2).astype(F)
    li = (rng.standard_normal((B, G, G, 3)) * 2).astype(F)
    t = lambda a: torch.from_numpy(a).cuda()
    pred = {"logits_areas": t(la), "logits_objects_exclusive": t(le), "logits_objects_independent": t(li),
            "bev_features": types.FeaturePlane(features=None, valid=t(bev_valid.astype(np.uint8)))}
    model = semantic_net.SemanticNetModel(cfg, gt_classes)
    losses, metrics = model.loss_metrics_function(pred, {"rasters": {"gt_semantics": masks}})
    torch.cuda.synchronize()
    ol, om = osn.loss_metrics(la, le, li, bev_valid, masks, gt_classes, cfg.area_classes, cfg.object_classes_exclusive,
                              cfg.object_classes_independent, dict(cfg.area_frequencies) if balanced else None,
                              dict(cfg.object_frequencies) if balanced else None)
    for k in ("nll_areas", "nll_objects_exclusive", "nll_objects_indep", "total"):
        got, ref = losses[k].cpu().numpy(), ol[k]
        assert np.abs(got - ref).max() <= 1e-4 * (1 + np.abs(ref).max()), (k, got, ref)
    assert losses["total"][2].item() == 0.0
    chk = lambda key, ref: np.testing.assert_allclose(metrics[f"semantics/{key}"].cpu().numpy(), ref, atol=1e-6)
    chk("accuracy", om["accuracy"]); chk("accuracy/excl", om["accuracy/excl"])
    chk("recall/average", om["recall/average"]); chk("recall/average/excl", om["recall/average/excl"])
    chk("recall/average/indep", om["recall/average/indep"])
    for i, n in enumerate(cfg.area_classes):
        chk(f"recall/{n}", om["recall_areas"][:, i])
    for i, n in enumerate((*cfg.object_classes_exclusive, "void")):
        chk(f"recall/{n}", om["recall_excl"][:, i])
    for i, n in enumerate(cfg.object_classes_independent):
        chk(f"recall/{n}", om["recall_indep"][:, i])


def test_label_preparation_kernel_equals_numpy():
    """snapb200_sem_labels vs the NumPy restatement of _create_exclusive_labels (semantic_net.py:254-298), 'line' merging
    included."""
    from oracle import semantic_net as osn
    from snap_b200 import ops
    rng = np.random.default_rng(4)
    gt = ("road", "sidewalk", "line", "stopline", "otherlanemarking", "tree", "pole", "traffic_sign", "street_light")
    area, excl, indep = ("sidewalk", "road", "line"), ("pole", "tree"), ("street_light", "traffic_sign")
    B, G = 2, 24
    masks = rng.random((B, G, G, len(gt))) < 0.15
    bev = rng.random((B, G, G)) < 0.8
    gi = {c: i for i, c in enumerate(gt)}
    sel = lambda cls: [[gi[n]] + ([gi[x] for x in ("stopline", "otherlanemarking") if x not in cls] if n == "line" else []) for n in cls]
    rows = B * G * G
    la = torch.empty(rows, dtype=torch.int32, device="cuda"); va = torch.empty(rows, dtype=torch.uint8, device="cuda")
    le = torch.empty(rows, dtype=torch.int32, device="cuda"); mi = torch.empty((rows, 2), dtype=torch.uint8, device="cuda")
    ops.sem_labels(sel(area), sel(excl), [gi[n] for n in indep], len(gt), torch.from_numpy(masks.view(np.uint8)).cuda().reshape(rows, -1),
                   torch.from_numpy(bev.astype(np.uint8)).cuda().reshape(-1), la, va, le, mi)
    ola, ova = osn.create_exclusive_labels(masks, gt, area)
    ole, _ = osn.create_exclusive_labels(masks, gt, excl, add_void=True)
    assert np.array_equal(la.cpu().numpy(), ola.reshape(-1)) and np.array_equal(va.cpu().numpy().astype(bool), (ova & bev).reshape(-1))
    assert np.array_equal(le.cpu().numpy(), ole.reshape(-1)) and (ole == 2).any()
    assert np.array_equal(mi.cpu().numpy().astype(bool), masks[..., [gi[n] for n in indep]].reshape(rows, 2))
