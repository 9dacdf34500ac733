"""`snap_b200/encoder_train.py::FPNBackward` on the emulated operator layer vs torch autograd of the oracle's FPN decoder
(`oracle/image_encoder.py::fpn_decoder`, pinned against the reference's own FPNDecoder): gradients of every skip_norm /
skip_conv array and the cotangents of the four skip inputs, from a cotangent on the finest level only."""
import numpy as np
import torch

from ops_emulation import emulated_ops, gn_stats
from util import F, bf16_np, rd_bf16


def test_fpn_backward_plan_matches_autograd():
    from oracle import image_encoder as oie
    from snap_b200 import configs, encoder_train, ops, params
    rng = np.random.default_rng(41)
    n, od = 2, 128
    shapes = [(2, 4, 512), (4, 8, 256), (8, 16, 128), (16, 32, 64)]          # coarse -> fine (a narrow trunk for speed)
    dec = {}
    for level, (h, w, c) in enumerate(shapes):
        dec[f"{level}_skip_norm"] = {"scale": (1 + 0.2 * rng.standard_normal((1, 1, 1, c))).astype(F),
                                     "bias": (0.1 * rng.standard_normal((1, 1, 1, c))).astype(F)}
        dec[f"{level}_skip_conv"] = {"kernel": (rng.standard_normal((1, 1, c, od)) / np.sqrt(c)).astype(F)}
    dec = params.round_to_bf16(dec)
    skips_np = [bf16_np(rng.standard_normal((n, h, w, c))) for h, w, c in shapes]
    dfin = bf16_np(rng.standard_normal((n, 16, 32, od)) * 0.1)
    # reference
    tp = {k: {a: torch.from_numpy(v).requires_grad_(True) for a, v in d.items()} for k, d in dec.items()}
    xs = [torch.from_numpy(s).requires_grad_(True) for s in skips_np]
    outs = oie.fpn_decoder(xs, tp, rd_bf16)
    (outs[-1] * torch.from_numpy(dfin)).sum().backward()
    # plan
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    with emulated_ops():
        fb = encoder_train.FPNBackward(dec, n, shapes, torch.device("cpu"), od)
        skips = [bf(s.reshape(-1, s.shape[-1])) for s in skips_np]
        accs = []
        for s, (h, w, c) in zip(skips, shapes):
            acc = torch.zeros((ops.GN_REPLICAS, n, 32, 2), dtype=torch.float64)
            gn_stats(s, n, h * w, c, True, acc)                                  # the forward's statistics of relu(skip)
            accs.append(acc)
        dsk = fb.backward(skips, accs, bf(dfin.reshape(-1, od)))
        got = fb.grads_tree()
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    errs = {}
    for k, d in tp.items():
        for a, t in d.items():
            assert float(t.grad.norm()) > 1e-6, (k, a)
            errs[f"{k}/{a}"] = rel(got[k][a], t.grad.numpy())
    for level, (x, (h, w, c)) in enumerate(zip(xs, shapes)):
        errs[f"skip{level}"] = rel(dsk[level][: n * h * w].float().numpy().reshape(n, h, w, c), x.grad.numpy())
    print({k: round(float(v), 4) for k, v in errs.items()})
    assert max(errs.values()) < 3e-2, errs


import pytest


@pytest.mark.parametrize("cin,nmid,nout,proj,n,G", [(64, 64, 256, True, 1, 8), (256, 64, 256, False, 2, 4),
                                                     (1024, 512, 2048, True, 1, 4)])
def test_bottleneck_unit_backward_plan_matches_autograd(cin, nmid, nout, proj, n, G):
    """`encoder_train.BottleneckUnitTrainer` (stride 1; identity or projection shortcut; weight gradients sliced for the
    trunk's wide layers) on the emulated operator layer vs autograd of the oracle's `residual_unit` (resnet.py:103-134)."""
    from oracle import resnet as ores
    from snap_b200 import encoder_train, ops, params
    rng = np.random.default_rng(cin + nout)
    ln = lambda *s: (rng.standard_normal(s) / np.sqrt(np.prod(s[:-1]))).astype(F)
    gnp = lambda c: {"scale": (1 + 0.2 * rng.standard_normal((1, 1, 1, c))).astype(F), "bias": (0.1 * rng.standard_normal((1, 1, 1, c))).astype(F)}
    p = {"gn1": gnp(cin), "gn2": gnp(nmid), "gn3": gnp(nmid), "conv1": {"kernel": ln(1, 1, cin, nmid)},
         "conv2": {"kernel": ln(3, 3, nmid, nmid)}, "conv3": {"kernel": ln(1, 1, nmid, nout)}}
    if proj:
        p["conv_proj"] = {"kernel": ln(1, 1, cin, nout)}
    p = params.round_to_bf16(p)
    x_np = bf16_np(rng.standard_normal((n, G, G, cin)))
    dout_np = bf16_np(rng.standard_normal((n, G, G, nout)) * 0.1)
    tp = {k: {a: torch.from_numpy(v).requires_grad_(True) for a, v in d.items()} for k, d in p.items()}
    xt = torch.from_numpy(x_np).requires_grad_(True)
    y = ores.residual_unit(xt, tp, 1, rd_bf16)
    (y * torch.from_numpy(dout_np)).sum().backward()
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    with emulated_ops():
        ut = encoder_train.BottleneckUnitTrainer(p, n, G, G, torch.device("cpu"))
        xb = bf(x_np.reshape(-1, cin))
        acc = torch.zeros((ops.GN_REPLICAS, n, 32, 2), dtype=torch.float64)
        gn_stats(xb, n, G * G, cin, False, acc)
        out = ut.forward(xb, acc)
        dx = ut.backward(bf(dout_np.reshape(-1, nout)))
        got = ut.grads_tree(p)
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    assert rel(out[: n * G * G].float().numpy().reshape(n, G, G, nout), y.detach().numpy()) < 2e-2
    errs = {"dx": rel(dx[: n * G * G].float().numpy().reshape(n, G, G, cin), xt.grad.numpy())}
    for k, d in tp.items():
        for a, t in d.items():
            errs[f"{k}/{a}"] = rel(got[k][a], t.grad.numpy())
    print({k: round(float(v), 4) for k, v in errs.items()})
    assert max(errs.values()) < 4e-2, errs


@pytest.mark.parametrize("cin,nmid,nout,n,G", [(256, 128, 512, 1, 16), (64, 64, 128, 2, 8)])
def test_strided_unit_backward_plan_matches_autograd(cin, nmid, nout, n, G):
    """`encoder_train.StridedUnitTrainer` (first unit of stages 2-4: stride-2 3x3 conv on the phase-split layout, strided
    projection shortcut) on the emulated operator layer vs autograd of the oracle's `residual_unit(stride=2)`."""
    from oracle import resnet as ores
    from snap_b200 import encoder_train, ops, params
    rng = np.random.default_rng(cin + nout + 1)
    ln = lambda *s: (rng.standard_normal(s) / np.sqrt(np.prod(s[:-1]))).astype(F)
    gnp = lambda c: {"scale": (1 + 0.2 * rng.standard_normal((1, 1, 1, c))).astype(F), "bias": (0.1 * rng.standard_normal((1, 1, 1, c))).astype(F)}
    p = params.round_to_bf16({"gn1": gnp(cin), "gn2": gnp(nmid), "gn3": gnp(nmid), "conv1": {"kernel": ln(1, 1, cin, nmid)},
                              "conv2": {"kernel": ln(3, 3, nmid, nmid)}, "conv3": {"kernel": ln(1, 1, nmid, nout)},
                              "conv_proj": {"kernel": ln(1, 1, cin, nout)}})
    x_np = bf16_np(rng.standard_normal((n, G, G, cin)))
    dout_np = bf16_np(rng.standard_normal((n, G // 2, G // 2, nout)) * 0.1)
    tp = {k: {a: torch.from_numpy(v).requires_grad_(True) for a, v in d.items()} for k, d in p.items()}
    xt = torch.from_numpy(x_np).requires_grad_(True)
    y = ores.residual_unit(xt, tp, 2, rd_bf16)
    assert y.shape == (n, G // 2, G // 2, nout)
    (y * torch.from_numpy(dout_np)).sum().backward()
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    with emulated_ops():
        ut = encoder_train.StridedUnitTrainer(p, n, G, G, torch.device("cpu"))
        xb = bf(x_np.reshape(-1, cin))
        acc = torch.zeros((ops.GN_REPLICAS, n, 32, 2), dtype=torch.float64)
        gn_stats(xb, n, G * G, cin, False, acc)
        out = ut.forward(xb, acc)
        dx = ut.backward(bf(dout_np.reshape(-1, nout)))
        got = ut.grads_tree(p)
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    ro = n * (G // 2) ** 2
    assert rel(out[:ro].float().numpy().reshape(y.shape), y.detach().numpy()) < 2e-2
    errs = {"dx": rel(dx[: n * G * G].float().numpy().reshape(n, G, G, cin), xt.grad.numpy())}
    for k, d in tp.items():
        for a, t in d.items():
            errs[f"{k}/{a}"] = rel(got[k][a], t.grad.numpy())
    print({k: round(float(v), 4) for k, v in errs.items()})
    assert max(errs.values()) < 4e-2, errs
