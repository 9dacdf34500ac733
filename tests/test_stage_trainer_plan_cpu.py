"""The launch plan of the 'resnet_stage' head training step (`snap_b200/semantic_train.py`) executed on the CPU with
the operator layer emulated in torch (`tests/ops_emulation.py`) against torch autograd of the oracle
(`oracle/semantic_net.py::stage_head_forward_torch`, whose forward is the NumPy `semantic_decoder` pinned against the
reference's own `SemanticNet.__call__`): buffers, offsets, operand layouts and launch order of forward and backward.
The CUDA kernels behind the operators are checked on the GPU (tests/test_stage_trainer_gpu.py)."""
import numpy as np
import torch

from ops_emulation import emulated_ops
from util import F, bf16_np, rd_bf16


def _setup(seed, B=2, G=8):
    from snap_b200 import configs, params
    rng = np.random.default_rng(seed)
    cfg = configs.semantic_net()
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_semantic_decoder(rng, cfg)))
    feats = bf16_np(rng.standard_normal((B, G, G, 128)) * 0.7)
    valid = rng.random((B, G, G)) < 0.8
    feats = feats * valid[..., None]
    return cfg, p, feats, valid, rng


def _torch_tree(tree, grad):
    return {k: (_torch_tree(v, grad) if isinstance(v, dict) else
                torch.from_numpy(np.ascontiguousarray(v, dtype=F)).requires_grad_(grad)) for k, v in tree.items()}


def _flat(tree, pre=()):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _flat(v, pre + (k,))
        else:
            yield pre + (k,), v


def test_torch_oracle_forward_equals_numpy_oracle():
    from oracle import semantic_net as osn
    cfg, p, feats, valid, _ = _setup(1)
    for rd in (lambda t: t, rd_bf16):
        ref = osn.semantic_decoder(feats, valid, p, rd)
        got = osn.stage_head_forward_torch(torch.from_numpy(feats), valid, _torch_tree(p, False), rd).numpy()
        assert got.shape == ref.shape == (2, 8, 8, 12)
        assert np.abs(got - ref).max() <= 1e-5 * (1 + np.abs(ref).max())


def test_stage_trainer_launch_plan_matches_autograd():
    from oracle import semantic_net as osn
    from snap_b200 import semantic_train, types
    cfg, p, feats, valid, rng = _setup(2)
    B, G = feats.shape[:2]
    plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16), torch.from_numpy(valid.astype(np.uint8)))
    Gmat = bf16_np(rng.standard_normal((B, G, G, 12)) * 0.05)
    with emulated_ops():
        tr = semantic_train.StageHeadTrainer(cfg, p, torch.device("cpu"))
        pred = tr.forward(plane)
        logits = torch.cat([pred["logits_areas"], pred["logits_objects_exclusive"], pred["logits_objects_independent"]], -1)
        buf = tr._buffers(B, G, G)
        buf["dlogits"].zero_()
        buf["dlogits"][: B * G * G, :12] = torch.from_numpy(Gmat * valid[..., None]).reshape(-1, 12).to(torch.bfloat16)
        tr.backward(plane, buf)
        grads, now = tr.grads_tree(), tr.params_tree()
    # forward: the plan reproduces the oracle in bf16-emulation mode
    tp = _torch_tree(p, True)
    ref_logits = osn.stage_head_forward_torch(torch.from_numpy(feats), valid, tp, rd_bf16)
    got = logits.numpy()
    assert not got[~valid].any()
    assert np.abs(got - ref_logits.detach().numpy()).max() <= 3e-2 * np.abs(ref_logits.detach().numpy()).max()
    # parameters round-trip through the trainer unchanged, with the Flax names and shapes
    for path, v in _flat(p):
        a = now
        for k in path:
            a = a[k]
        assert a.shape == np.asarray(v).shape and np.array_equal(a, np.asarray(v, dtype=F)), path
    # backward: every parameter gradient vs autograd of loss = sum(logits * G)
    (ref_logits * torch.from_numpy(Gmat)).sum().backward()
    worst = 0.0
    for path, t in _flat(tp):
        g = grads
        for k in path:
            g = g[k]
        r = t.grad.numpy()
        assert g.shape == r.shape, path
        err = np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
        worst = max(worst, err)
        assert err < 3e-2, ("/".join(path), err, np.linalg.norm(r))
    print(f"worst relative gradient error over {len(list(_flat(tp)))} parameter arrays: {worst:.4f}")


def _dp_worker(rank, world, port, out):
    """One gloo rank: the stage trainer's backward on this rank's half of the batch + the gradient mean over ranks."""
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snap_b200 import parallel, semantic_train, types
    cfg, p, feats, valid, rng = _setup(7, B=2)
    Gm = bf16_np(np.random.default_rng(8).standard_normal(feats.shape[:3] + (12,)) * 0.05)
    sl = slice(rank, rank + 1)                                                   # per-GPU batch 1 (trainer.py:452-464)
    with emulated_ops():
        tr = semantic_train.StageHeadTrainer(cfg, p, torch.device("cpu"))
        plane = types.FeaturePlane(torch.from_numpy(feats[sl]).to(torch.bfloat16), torch.from_numpy(valid[sl].astype(np.uint8)))
        tr.forward(plane)
        buf = tr._buffers(1, feats.shape[1], feats.shape[2])
        buf["dlogits"].zero_()
        buf["dlogits"][: feats.shape[1] * feats.shape[2], :12] = torch.from_numpy(
            (Gm[sl] * valid[sl][..., None]).reshape(-1, 12)).to(torch.bfloat16)   # d mean over the LOCAL batch (of 1) / d logits
        tr.backward(plane, buf)
        calls = parallel.pmean_tree({"/".join(r["path"]): r["g"] for r in tr.rec})      # as train_step does
        grads = tr.grads_tree()
    if rank == 0:
        torch.save({"grads": grads, "calls": calls}, out)
    dist.destroy_process_group()


def test_stage_trainer_gradient_mean_world_size_2(tmp_path):
    """Two gloo ranks with one scene each: mean over ranks of the shard gradients (one bucketed all-reduce) == the
    gradient of the whole batch on one process (GroupNorm statistics are per image, so the split is exact)."""
    import socket
    import torch.multiprocessing as mp
    from snap_b200 import semantic_train, types
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    assert got["calls"] == 1
    cfg, p, feats, valid, rng = _setup(7, B=2)
    Gm = bf16_np(np.random.default_rng(8).standard_normal(feats.shape[:3] + (12,)) * 0.05)
    B, G = feats.shape[:2]
    with emulated_ops():
        tr = semantic_train.StageHeadTrainer(cfg, p, torch.device("cpu"))
        plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16), torch.from_numpy(valid.astype(np.uint8)))
        tr.forward(plane)
        buf = tr._buffers(B, G, G)
        buf["dlogits"].zero_()
        # whole batch on one process: the loss is the mean over the B examples (trainer.py:221)
        buf["dlogits"][: B * G * G, :12] = torch.from_numpy((Gm * valid[..., None]).reshape(-1, 12) / B).to(torch.bfloat16)
        tr.backward(plane, buf)
        ref = tr.grads_tree()
    for path, r in _flat(ref):
        g = got["grads"]
        for k in path:
            g = g[k]
        err = np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)               # pmean of per-rank means = global mean
        assert err < 2e-2, ("/".join(path), err)


GT = ("road", "crosswalk", "sidewalk", "terrain", "building", "fence", "pole", "tree", "traffic_sign", "traffic_light",
      "street_light")


def test_stage_trainer_whole_train_step_on_the_emulated_layer():
    """`StageHeadTrainer.train_step` end to end (label preparation, balanced losses, loss gradient, backward, Adam) with
    every operator emulated: the gradients equal autograd of the oracle's loss through the oracle's decoder, and a few
    steps reduce the loss.  (The loss kernels themselves are GPU-verified, tests/test_semantic_gpu.py.)"""
    from oracle import semantic_net as osn
    from snap_b200 import semantic_net, semantic_train, types
    cfg, p, feats, valid, rng = _setup(11, B=2, G=8)
    cfg.area_frequencies = tuple(zip(cfg.area_classes, (0.036434, 0.226553, 0.446990, 0.085374, 0.204649)))
    cfg.object_frequencies = (("fence", 0.006257), ("pole", 0.001172), ("tree", 0.001924), ("traffic_sign", 0.000960),
                              ("traffic_light", 0.000559), ("street_light", 0.000738), ("void", 0.988391))
    masks = rng.random(feats.shape[:3] + (len(GT),)) < 0.25
    plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16), torch.from_numpy(valid.astype(np.uint8)))
    data = {"rasters": {"gt_semantics": masks}}
    with emulated_ops():
        model = semantic_net.SemanticNetModel(cfg, GT)
        tr = semantic_train.StageHeadTrainer(cfg, p, torch.device("cpu"), lr=3e-3)
        total, _, _ = tr.train_step(plane, model, data, update=False)
        grads = tr.grads_tree()
        hist = []
        for _ in range(6):
            hist.append(float(tr.train_step(plane, model, data)[0].mean()))
        new = tr.params_tree()
    # reference gradients: oracle decoder + oracle loss
    tp = _torch_tree(p, True)
    logits = osn.stage_head_forward_torch(torch.from_numpy(feats), valid, tp, rd_bf16)
    la, va = osn.create_exclusive_labels(masks, GT, cfg.area_classes)
    le, _ = osn.create_exclusive_labels(masks, GT, cfg.object_classes_exclusive, add_void=True)
    gi = {c: i for i, c in enumerate(GT)}
    mi = masks[..., [gi[c] for c in cfg.object_classes_independent]]
    fa, fo = dict(cfg.area_frequencies), dict(cfg.object_frequencies)
    w = (osn.balancing_weights(fa, cfg.area_classes), osn.balancing_weights(fo, (*cfg.object_classes_exclusive, "void")),
         *osn.balancing_weights(fo, cfg.object_classes_independent, binary=True))
    loss, ref_total = osn.total_loss_torch(logits, la, va, le, mi, valid, 5, 4, *w)
    loss.backward()
    assert np.abs(total.numpy() - ref_total.detach().numpy()).max() <= 3e-2 * (1 + np.abs(ref_total.detach().numpy()).max())
    for path, t in _flat(tp):
        g = grads
        for k in path:
            g = g[k]
        err = np.linalg.norm(g - t.grad.numpy()) / (np.linalg.norm(t.grad.numpy()) + 1e-30)
        assert err < 5e-2, ("/".join(path), err)
    print("loss:", " ".join(f"{h:.4f}" for h in hist))
    assert np.isfinite(hist).all() and hist[-1] < hist[0]
    assert new["layers_3"]["Dense_1"]["kernel"].shape == (256, 12)
    assert not np.array_equal(new["layers_1"]["unit01"]["gn2"]["scale"], p["layers_1"]["unit01"]["gn2"]["scale"])
