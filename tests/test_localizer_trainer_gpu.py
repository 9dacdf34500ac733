"""`snap_b200.localizer_trainer.LocalizerTrainer`: the config-4 training step (`snap/trainer.py:165-295` around
`snap/models/bev_localizer.py`) with frozen image encoders, at BASELINE configs[3] size (G = 128, 4 x 640 x 480 map views,
4,652 field-of-view query points, 10,000 pose samples)."""
import numpy as np
import pytest
import torch

from util import F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]


def _setup(batch=2, seed=9):
    from snap_b200 import bev_localizer, configs, params, synthetic, types
    G, hw = 128, (480, 640)
    rng = np.random.default_rng(seed)
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview",))
    cfg.filter_points_in_fov = True
    cfg.num_pose_samples = 2_000
    cfg.num_pose_sampling_retries = 2
    grid = types.Grid2D((G, G), 0.2)
    loc = bev_localizer.BEVLocalizer(cfg, None, grid)
    mp = params.round_to_bf16(params.perturb_affine(rng, params.init_bev_mapper(rng, cfg.bev_mapper)))
    p = loc.init_params(mp)
    data = synthetic.make_tile(61, 4, hw, G, batch=batch)
    v = 2
    T, cam = data["T_view2scene"], data["camera"]
    t_q2m = (np.round(T.t[:, v, :2] / 0.2) * 0.2).astype(F)                       # [B, 2]
    z_off = (np.median(T.t[..., -1].astype(F), axis=-1).astype(F) - F(4.0)).astype(F)
    shift = np.concatenate([t_q2m, np.zeros((batch, 1), F)], -1)[:, None]
    query = {"images": np.ascontiguousarray(data["images"][:, [v]]),
             "camera": types.Camera(wh=cam.wh[:, [v]].copy(), f=cam.f[:, [v]].copy(), c=cam.c[:, [v]].copy()),
             "T_view2scene": types.Transform3D(R=T.R[:, [v]].copy(), t=(T.t[:, [v]] - shift).astype(F)), "z_offset": z_off}
    T_q2m = types.Transform3D(R=np.broadcast_to(np.eye(3, dtype=F), (batch, 3, 3)).copy(),
                              t=np.concatenate([t_q2m, np.zeros((batch, 1), F)], -1))
    batch_data = {"map": {**data, "z_offset": z_off}, "query": query, "T_query2map": T_q2m}
    return loc, p, batch_data


def _gen(seed=1):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def test_training_step_reduces_the_loss_and_refreshes_the_forward_weights():
    from snap_b200 import localizer_trainer
    loc, p, data = _setup()
    tr = localizer_trainer.LocalizerTrainer(loc, p, lr=2e-3)
    enc_tree = p["bev_mapper"]["streetview_encoder"]["image_encoder"]
    losses = []
    for it in range(6):
        total, _, metrics = tr.train_step(data, {"sampling": _gen()})
        torch.cuda.synchronize()
        losses.append(float(total.mean()))
        assert metrics["is_finite"] and np.isfinite(metrics["l2_grads"]) and metrics["l2_grads"] > 0
        print(f"step {it}: nll {losses[-1]:.4f}  |g| {metrics['l2_grads']:.4e}  T {float(tr.params['temperature']):.4f}")
    assert tr.step == 6 and tr.skipped_steps == 0
    assert losses[-1] < losses[0] - 1e-3, losses
    # the frozen encoder tree kept its identity (cached launch plan), the trainable sub-trees are new objects with new values
    new = tr.params
    assert new["bev_mapper"]["streetview_encoder"]["image_encoder"] is enc_tree
    assert new["bev_mapper"]["matching_proj"] is not p["bev_mapper"]["matching_proj"]
    for a, b in ((new["bev_mapper"]["matching_proj"]["kernel"], p["bev_mapper"]["matching_proj"]["kernel"]),
                 (new["bev_mapper"]["streetview_encoder"]["fusion_mlp"]["Dense_1"]["kernel"],
                  p["bev_mapper"]["streetview_encoder"]["fusion_mlp"]["Dense_1"]["kernel"])):
        assert a.shape == b.shape and np.abs(a - b).max() > 0
    assert float(new["temperature"]) != float(p["temperature"])
    # the encoder-feature cotangents are there for `TrunkTrainer.backward` (one per scene and side)
    pred, _, _ = tr.loss_and_gradients(data, {"sampling": _gen()})
    enc = pred["encoder_cotangents"]
    assert len(enc["map"]) == 2 and len(enc["query"]) == 2
    assert enc["map"][0].shape[1] == 128 and torch.isfinite(enc["map"][0].float()).all() and enc["map"][0].float().abs().max() > 0


def test_temperature_gradient_matches_a_finite_difference_and_nan_steps_are_skipped():
    from snap_b200 import localizer_trainer
    loc, p, data = _setup(batch=1)
    tr = localizer_trainer.LocalizerTrainer(loc, p, lr=1e-3, max_grad_norm=0.5)
    total, _, metrics = tr.train_step(data, {"sampling": _gen()}, update=False)
    g_T = float(tr.grads_tree()["temperature"])
    # the sampled poses do not depend on the temperature's value beyond the soft-max they are drawn from: score the SAME
    # poses at T +- eps through the forward's own kernels
    from snap_b200 import pose_estimation
    pred = loc.apply({"params": p}, data, rngs={"sampling": _gen()})
    poses = pred["map_t_query_samples"]
    pq, pm = pred["query"]["bev_matching"], pred["map"]["bev_matching"]
    q_xy = torch.from_numpy(np.ascontiguousarray(loc.q_xy_p[:, 0])).cuda()

    def nll_at(T):
        maps = pose_estimation.point_similarities(pq.features.reshape(1, -1, 32), pq.valid.reshape(1, -1), pm.features, T, True, None)
        sc = pose_estimation.pose_scoring_many_batched(poses, maps, q_xy, pm.valid, loc.grid_map, False).double()
        return float(-(sc[0, 0] - torch.logsumexp(sc[0], 0)))
    T0, eps = float(p["temperature"]), 0.05
    fd = (nll_at(T0 + eps) - nll_at(T0 - eps)) / (2 * eps)
    print(f"d nll / d temperature: backward {g_T:.5f}, finite difference {fd:.5f}; nll {float(total.mean()):.4f}")
    assert abs(g_T - fd) <= 0.05 * abs(fd) + 1e-3
    # clipping: the applied gradient has norm <= max_grad_norm
    before = tr.params
    tr.bucket.flat.mul_(100.0)
    assert tr.apply_update() and tr.params is not before
    # a non-finite gradient skips the update: parameters, optimiser state and step count stay
    snap = (tr.params, tr.masters.flat.clone(), tr.mom[0].clone(), tr.step)
    tr.bucket.flat[5] = float("nan")
    assert tr.apply_update() is False and tr.skipped_steps == 1
    assert tr.params is snap[0] and torch.equal(tr.masters.flat, snap[1]) and torch.equal(tr.mom[0], snap[2]) and tr.step == snap[3]


def test_full_training_step_with_the_image_encoder():
    """`train_encoder=True`: the street-view encoder's training forward / backward (`encoder_train.TrunkTrainer` over the map
    and query images) inside the step; all 23.5 M encoder gradients are finite and non-zero, the masters move, the loss on
    the fixed batch goes down, and the un-trained forward (EncoderPlan) agrees with the training forward's first loss."""
    from snap_b200 import localizer_trainer
    loc, p, data = _setup(batch=1)
    frozen = localizer_trainer.LocalizerTrainer(loc, p, lr=1e-3)
    l_frozen = float(frozen.train_step(data, {"sampling": _gen()}, update=False)[0].mean())
    tr = localizer_trainer.LocalizerTrainer(loc, p, lr=3e-4, train_encoder=True)
    losses = []
    for it in range(5):
        total, _, metrics = tr.train_step(data, {"sampling": _gen()})
        torch.cuda.synchronize()
        losses.append(float(total.mean()))
        assert metrics["is_finite"] and np.isfinite(metrics["l2_grads"])
        if it == 0:
            before = [pt.clone() for _, pt, _ in tr.enc.leaves[:3]]
            n_par = sum(pt.numel() for _, pt, _ in tr.enc.leaves)
            nz = [float(g.abs().max()) > 0 for g in tr.enc_bucket.views]
            print(f"encoder leaves {len(tr.enc.leaves)}, parameters {n_par}, non-zero gradients {sum(nz)}/{len(nz)}")
            assert n_par > 23_000_000 and all(nz) and tr.enc_bucket.nbytes > 90_000_000
        print(f"step {it}: nll {losses[-1]:.4f}  |g| {metrics['l2_grads']:.4e}")
    # the training forward (im2col root conv, per-unit buffers) and the inference plan are two launch sequences of the same
    # network: the first loss agrees up to bf16 noise of the free-running encoder
    assert abs(losses[0] - l_frozen) < 0.15 * abs(l_frozen) + 0.1, (losses[0], l_frozen)
    assert losses[-1] < losses[0] - 1e-3, losses
    tree = tr.encoder_params_tree()
    k0 = np.asarray(p["bev_mapper"]["streetview_encoder"]["image_encoder"]["encoder"]["root_block"]["conv_root"]["kernel"])
    assert tree["encoder"]["root_block"]["conv_root"]["kernel"].shape == k0.shape
    assert np.abs(tree["encoder"]["root_block"]["conv_root"]["kernel"] - k0).max() > 0


def test_batch_mask_weights_the_examples_like_a_masked_mean():
    """`trainer.py:221`: loss = mean over batch['batch_mask'].  On one batch and one sampling seed the gradients are linear in
    the example weights: g(mask [1,0]) + g(mask [0,1]) = 2 g(mask [1,1]), and an all-ones mask equals no mask."""
    from snap_b200 import localizer_trainer
    loc, p, data = _setup(batch=2)
    tr = localizer_trainer.LocalizerTrainer(loc, p, lr=1e-3)

    def grads(mask):
        d = dict(data) if mask is None else dict(data, batch_mask=np.asarray(mask))
        tr.train_step(d, {"sampling": _gen()}, update=False)
        torch.cuda.synchronize()
        return tr.bucket.flat.clone()
    g_none, g11, g10, g01 = grads(None), grads([1, 1]), grads([1, 0]), grads([0, 1])
    rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
    print(f"all-ones mask vs none {rel(g11, g_none):.2e}; g10 + g01 vs 2 g11 {rel(g10 + g01, 2 * g11):.2e}; |g10| {float(g10.norm()):.3e} |g01| {float(g01.norm()):.3e}")
    assert rel(g11, g_none) < 1e-3            # fp32 atomics in the scatter-add: order-dependent last bits
    assert rel(g10 + g01, 2 * g11) < 2e-2     # bf16 cotangent rows round differently at different scales
    assert rel(g10, g01) > 1e-2               # the two examples do differ
