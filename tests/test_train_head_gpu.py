"""Head-only training step of the semantic fine-tuning configuration (default 'mlp' decoder) on the GPU against torch
autograd of the oracle restatement (oracle/semantic_net.py: mlp_head_forward_torch + total_loss_torch)."""
import numpy as np
import pytest
import torch

from util import F, bf16_np

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

GT = ("road", "crosswalk", "sidewalk", "terrain", "building", "fence", "pole", "tree", "traffic_sign", "traffic_light",
      "street_light")


def _setup(seed, B=2, G=32, balanced=True):
    from snap_b200 import configs, params, semantic_net, types
    rng = np.random.default_rng(seed)
    cfg = configs.semantic_net()
    cfg.decoder_type, cfg.decoder_dim, cfg.mlp_num_layers = "mlp", 128, 2          # defaults.py:289-291
    if balanced:
        cfg.area_frequencies = tuple(zip(cfg.area_classes, (0.036434, 0.226553, 0.446990, 0.085374, 0.204649)))
        cfg.object_frequencies = (("fence", 0.006257), ("pole", 0.001172), ("tree", 0.001924), ("traffic_sign", 0.000960),
                                  ("traffic_light", 0.000559), ("street_light", 0.000738), ("void", 0.988391))
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_mlp(rng, 128, (128, 128, 12))))
    feats = bf16_np(rng.standard_normal((B, G, G, 128)) * 0.7)
    valid = rng.random((B, G, G)) < 0.75
    feats = feats * valid[..., None]
    masks = rng.random((B, G, G, len(GT))) < 0.25
    plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16).cuda(), torch.from_numpy(valid.astype(np.uint8)).cuda())
    model = semantic_net.SemanticNetModel(cfg, GT)
    return cfg, p, feats, valid, masks, plane, model


def _oracle_grads(cfg, p, feats, valid, masks, balanced):
    from oracle import semantic_net as osn
    tp = {k: {n: torch.from_numpy(np.ascontiguousarray(v[n], dtype=F)).requires_grad_(True) for n in ("kernel", "bias")}
          for k, v in p.items()}
    rd = lambda t: t.to(torch.bfloat16).float()
    logits = osn.mlp_head_forward_torch(torch.from_numpy(feats), valid, tp, rd)
    la, va = osn.create_exclusive_labels(masks, GT, cfg.area_classes)
    le, _ = osn.create_exclusive_labels(masks, GT, cfg.object_classes_exclusive, add_void=True)
    gi = {c: i for i, c in enumerate(GT)}
    mi = masks[..., [gi[c] for c in cfg.object_classes_independent]]
    w = (None,) * 4
    if balanced:
        fa, fo = dict(cfg.area_frequencies), dict(cfg.object_frequencies)
        w = (osn.balancing_weights(fa, cfg.area_classes), osn.balancing_weights(fo, (*cfg.object_classes_exclusive, "void")),
             *osn.balancing_weights(fo, cfg.object_classes_independent, binary=True))
    loss, total = osn.total_loss_torch(logits, la, va, le, mi, valid, 5, 4, *w)
    loss.backward()
    return float(loss), total.detach().numpy(), {k: {n: t.grad.numpy() for n, t in v.items()} for k, v in tp.items()}, \
        logits.detach().numpy()


@pytest.mark.parametrize("balanced", [True, False])
def test_gradients_match_autograd(balanced):
    from snap_b200 import semantic_net
    cfg, p, feats, valid, masks, plane, model = _setup(21, balanced=balanced)
    tr = semantic_net.MLPHeadTrainer(cfg, p, plane.features.device)
    total, losses, metrics = tr.train_step(plane, model, {"rasters": {"gt_semantics": masks}}, update=False)
    torch.cuda.synchronize()
    ref_loss, ref_total, ref_g, ref_logits = _oracle_grads(cfg, p, feats, valid, masks, balanced)
    got_total = total.cpu().numpy()
    assert np.abs(got_total - ref_total).max() <= 2e-3 * (1 + np.abs(ref_total).max()), (got_total, ref_total)
    for i, n in enumerate(tr.names):
        cout = tr.dims[i][2]
        gW, gb = tr.dW[i][:, :cout].cpu().numpy(), tr.db[i][:cout].cpu().numpy()
        rW, rb = ref_g[n]["kernel"], ref_g[n]["bias"]
        eW = np.linalg.norm(gW - rW) / (np.linalg.norm(rW) + 1e-30)
        eb = np.linalg.norm(gb - rb) / (np.linalg.norm(rb) + 1e-30)
        print(f"balanced={balanced} {n}: |dW| {np.linalg.norm(rW):.4e} rel err {eW:.4f}, |db| {np.linalg.norm(rb):.4e} rel err {eb:.4f}")
        # bf16 activations / cotangents on the GPU vs fp32 cotangents in autograd: a few 1e-3 relative
        assert eW < 2e-2 and eb < 2e-2
        assert not tr.dW[i][:, cout:].any() and not tr.db[i][cout:].any(), "padding columns carry no gradient"


def test_wgrad_relu_adam_kernels_exact_cases():
    from snap_b200 import ops
    rng = np.random.default_rng(3)
    M, K, N = 4096 + 48, 96, 48
    x = bf16_np(rng.standard_normal((M, K)))
    dy = bf16_np(rng.standard_normal((M, N)) * 0.1)
    xd, dyd = torch.from_numpy(x).to(torch.bfloat16).cuda(), torch.from_numpy(dy).to(torch.bfloat16).cuda()
    dW = torch.empty((K, N), dtype=torch.float32, device="cuda")
    db = torch.empty((N,), dtype=torch.float32, device="cuda")
    ops.dense_wgrad(xd, dyd, M, K, N, dW, db)
    ref = x.astype(np.float64).T @ dy.astype(np.float64)
    assert np.abs(dW.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.abs(db.cpu().numpy() - dy.astype(np.float64).sum(0)).max() <= 1e-4 * np.abs(dy.sum(0)).max() + 1e-5
    # relu backward: exact masking
    h = bf16_np(np.maximum(rng.standard_normal((256, 64)), 0))
    d = bf16_np(rng.standard_normal((256, 64)))
    dd = torch.from_numpy(d).to(torch.bfloat16).cuda()
    ops.relu_bwd(torch.from_numpy(h).to(torch.bfloat16).cuda(), dd, h.size)
    assert np.array_equal(dd.float().cpu().numpy(), np.where(h > 0, d, 0))
    # adam: two steps vs the optax.adam recurrences in float64
    n = 1000
    p0, g1, g2 = rng.standard_normal(n).astype(F), rng.standard_normal(n).astype(F), rng.standard_normal(n).astype(F)
    pt, m, v = torch.from_numpy(p0.copy()).cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pr, mr, vr = p0.astype(np.float64), np.zeros(n), np.zeros(n)
    for t, g in enumerate((g1, g2), 1):
        ops.adam_step(pt, m, v, torch.from_numpy(g).cuda(), 1e-2, t)
        mr = 0.9 * mr + 0.1 * g
        vr = 0.999 * vr + 0.001 * g.astype(np.float64) ** 2
        pr = pr - 1e-2 * (mr / (1 - 0.9 ** t)) / (np.sqrt(vr / (1 - 0.999 ** t)) + 1e-8)
    assert np.abs(pt.cpu().numpy() - pr).max() <= 1e-5


def test_training_reduces_the_loss_and_forward_matches_oracle():
    from oracle import semantic_net as osn
    from snap_b200 import semantic_net
    cfg, p, feats, valid, masks, plane, model = _setup(22, balanced=True)
    head = semantic_net.SemanticHead(cfg)
    pred = head.apply({"params": {"decoder": p}}, plane)
    logits = torch.cat([pred["logits_areas"], pred["logits_objects_exclusive"], pred["logits_objects_independent"]], -1)
    tp = {k: {n: torch.from_numpy(np.ascontiguousarray(v[n], dtype=F)) for n in ("kernel", "bias")} for k, v in p.items()}
    ref = osn.mlp_head_forward_torch(torch.from_numpy(feats), valid, tp, lambda t: t.to(torch.bfloat16).float()).numpy()
    got = logits.cpu().numpy()
    assert got.shape == ref.shape == (2, 32, 32, 12)
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max() and not got[~valid].any()
    tr = semantic_net.MLPHeadTrainer(cfg, p, plane.features.device, lr=1e-2)
    data = {"rasters": {"gt_semantics": masks}}
    hist = []
    for _ in range(30):
        total, _, _ = tr.train_step(plane, model, data)
        hist.append(float(total.mean().item()))
    print("loss:", " ".join(f"{h:.4f}" for h in hist[::4]))
    assert hist[-1] < 0.9 * hist[0] and np.isfinite(hist).all()
    new = tr.params_tree()
    assert new["Dense_2"]["kernel"].shape == (128, 12) and not np.array_equal(new["Dense_0"]["kernel"], p["Dense_0"]["kernel"])
