"""Design checks for the round-2 backward kernels (tools/design/backward_formulas.py) against torch autograd on the CPU:
GroupNorm and StdConv closed forms, and the 3x3 conv backward as segment GEMMs / shifted split-K products on the
zero-bordered layout of the forward engine."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "design"))
import backward_formulas as bf  # noqa: E402


def test_groupnorm_backward_closed_form():
    rng = np.random.default_rng(0)
    N, H, W, C = 2, 5, 6, 64
    x = torch.tensor(rng.standard_normal((N, H, W, C)), dtype=torch.float64, requires_grad=True)
    scale = torch.tensor(rng.standard_normal(C) * 0.3 + 1, dtype=torch.float64, requires_grad=True)
    bias = torch.tensor(rng.standard_normal(C) * 0.1, dtype=torch.float64, requires_grad=True)
    xg = x.reshape(N, H * W, 32, C // 32)
    mu = xg.mean(dim=(1, 3), keepdim=True)
    var = ((xg - mu) ** 2).mean(dim=(1, 3), keepdim=True)
    y = ((xg - mu) / torch.sqrt(var + 1e-5)).reshape(N, H, W, C) * scale + bias      # resnet.py:34-41,57-69
    dy = torch.tensor(rng.standard_normal((N, H, W, C)), dtype=torch.float64)
    y.backward(dy)
    dx, dscale, dbias = bf.groupnorm_backward(x.detach().numpy(), dy.numpy(), scale.detach().numpy())
    assert np.abs(dx - x.grad.numpy()).max() < 1e-10
    assert np.abs(dscale - scale.grad.numpy()).max() < 1e-10 and np.abs(dbias - bias.grad.numpy()).max() < 1e-10


def test_stdconv_weight_backward_closed_form():
    rng = np.random.default_rng(1)
    w = torch.tensor(rng.standard_normal((3, 3, 8, 5)), dtype=torch.float64, requires_grad=True)
    mu = w.mean(dim=(0, 1, 2), keepdim=True)
    var = ((w - mu) ** 2).mean(dim=(0, 1, 2), keepdim=True)
    ws = (w - mu) / torch.sqrt(var + 1e-10)                                          # resnet.py:73-79
    dws = torch.tensor(rng.standard_normal((3, 3, 8, 5)), dtype=torch.float64)
    ws.backward(dws)
    assert np.abs(bf.stdconv_weight_backward(w.detach().numpy(), dws.numpy()) - w.grad.numpy()).max() < 1e-10


def test_conv3x3_backward_on_the_bordered_layout():
    rng = np.random.default_rng(2)
    N, H, W, Cin, Cout = 2, 6, 7, 4, 3
    x = torch.tensor(rng.standard_normal((N, H, W, Cin)), dtype=torch.float64, requires_grad=True)
    w = torch.tensor(rng.standard_normal((3, 3, Cin, Cout)), dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    # the forward decomposition the engine already runs
    assert np.abs(bf.conv3x3_forward_segments(x.detach().numpy(), w.detach().numpy()) - y.detach().numpy()).max() < 1e-10
    dy = torch.tensor(rng.standard_normal((N, H, W, Cout)), dtype=torch.float64)
    y.backward(dy)
    dx = bf.conv3x3_dx_segments(dy.numpy(), w.detach().numpy())
    dw = bf.conv3x3_dw_shifted(x.detach().numpy(), dy.numpy())
    assert np.abs(dx - x.grad.numpy()).max() < 1e-10
    assert np.abs(dw - w.grad.numpy()).max() < 1e-10


def test_multiview_pooling_backward_closed_form():
    rng = np.random.default_rng(3)
    for valid in ([True, True, False, True], [False, True, False, False]):
        V, D = 4, 6
        valid = np.array(valid)
        f = torch.tensor(rng.standard_normal((V, D)), dtype=torch.float64, requires_grad=True)
        s = torch.tensor(rng.standard_normal(V) * 2, dtype=torch.float64, requires_grad=True)
        vm = torch.tensor(valid)
        mx = torch.clamp(torch.where(vm, s, torch.tensor(-np.inf, dtype=torch.float64)).max(), min=0.0).detach()
        e = torch.where(vm, torch.exp(s - mx), torch.zeros((), dtype=torch.float64))
        w = e / e.sum()
        mean = (w[:, None] * f).sum(0)
        var = (w[:, None] * (f - mean) ** 2).sum(0)
        smax = torch.where(vm, s, torch.tensor(-np.inf, dtype=torch.float64)).max()
        dmean, dvar, dsmax = rng.standard_normal(D), rng.standard_normal(D), float(rng.standard_normal())
        ((mean * torch.tensor(dmean)).sum() + (var * torch.tensor(dvar)).sum() + smax * dsmax).backward()
        df, ds = bf.pool_multiview_backward(f.detach().numpy(), s.detach().numpy(), valid, dmean, dvar, dsmax)
        assert np.abs(df - f.grad.numpy()).max() < 1e-10 and np.abs(ds - s.grad.numpy()).max() < 1e-10


def test_correlation_backward_as_two_correlations():
    rng = np.random.default_rng(4)
    R, G, D = 3, 5, 4
    q = torch.tensor(rng.standard_normal((R, G, G, D)), dtype=torch.float64, requires_grad=True)
    m = torch.tensor(rng.standard_normal((G, G, D)), dtype=torch.float64, requires_grad=True)
    m_pad = torch.nn.functional.pad(m.permute(2, 0, 1)[None], (G - 1, G - 1, G - 1, G - 1), mode="replicate")
    S = torch.nn.functional.conv2d(m_pad, q.permute(0, 3, 1, 2))[0]                  # [R, 2G-1, 2G-1]
    dS = torch.tensor(rng.standard_normal(S.shape), dtype=torch.float64)
    S.backward(dS)
    dq, dm = bf.xcorr_backward(q.detach().numpy(), m.detach().numpy(), dS.numpy())
    assert np.abs(dq - q.grad.numpy()).max() < 1e-10 and np.abs(dm - m.grad.numpy()).max() < 1e-10


def test_residual_unit_backward_composition():
    """The whole pre-activation bottleneck unit (resnet.py:103-134) backward, composed of the building blocks in the launch
    order planned for round 2, against torch autograd of the oracle's residual_unit."""
    from oracle import resnet as ores
    rng = np.random.default_rng(5)
    N, H, W, C, M = 2, 5, 6, 128, 32
    p = {g: {"scale": 1 + 0.2 * rng.standard_normal((1, 1, 1, c)), "bias": 0.2 * rng.standard_normal((1, 1, 1, c))}
         for g, c in (("gn1", C), ("gn2", M), ("gn3", M))}
    p.update(conv1={"kernel": rng.standard_normal((1, 1, C, M)) * 0.1}, conv2={"kernel": rng.standard_normal((3, 3, M, M)) * 0.1},
             conv3={"kernel": rng.standard_normal((1, 1, M, C)) * 0.1})
    x = rng.standard_normal((N, H, W, C))
    dy = rng.standard_normal((N, H, W, C))
    tp = {k: {n: torch.tensor(v, dtype=torch.float64, requires_grad=True) for n, v in d.items()} for k, d in p.items()}
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    # the oracle casts to fp32 inside standardize: run it in float64 by patching .float() away
    orig = ores.standardize
    ores.standardize = lambda t, dims, eps: (t - t.mean(dim=dims, keepdim=True)) / torch.sqrt(((t - t.mean(dim=dims, keepdim=True)) ** 2).mean(dim=dims, keepdim=True) + eps)
    try:
        ty = ores.residual_unit(tx, tp, 1)
    finally:
        ores.standardize = orig
    ty.backward(torch.tensor(dy))
    y, saved = bf.residual_unit_forward(x, p)
    assert np.abs(y - ty.detach().numpy()).max() < 1e-9
    dx, g = bf.residual_unit_backward(dy, saved, p)
    assert np.abs(dx - tx.grad.numpy()).max() < 1e-8
    for k, d in tp.items():
        for n, t in d.items():
            ref = t.grad.numpy().reshape(g[f"{k}/{n}"].shape)
            assert np.abs(g[f"{k}/{n}"] - ref).max() < 1e-8 * (1 + np.abs(ref).max()), (k, n)


def test_lift_gather_backward_scatter():
    """Bilinear gather + depth-score interpolation of one (voxel, view) pair: the scatter-add backward vs autograd, including a
    point at the image border whose clamped taps coincide."""
    rng = np.random.default_rng(6)
    Hf, Wf, D, S = 6, 7, 5, 8
    for p2d, depth in (((2.3, 4.6), 5.0), ((0.2, 6.9), 20.0)):          # (row, col) in texels; second: clamped taps
        fimg = torch.tensor(rng.standard_normal((Hf, Wf, D + S)), dtype=torch.float64, requires_grad=True)
        cr, cc = p2d[0] - 0.5, p2d[1] - 0.5                             # grids.py:129
        r0, c0 = int(np.floor(cr)), int(np.floor(cc))
        wr1, wc1 = cr - r0, cc - c0
        taps = [(min(max(r0 + i, 0), Hf - 1), min(max(c0 + j, 0), Wf - 1)) for i in (0, 1) for j in (0, 1)]
        weights = [(wr1 if i else 1 - wr1) * (wc1 if j else 1 - wc1) for i in (0, 1) for j in (0, 1)]
        t = np.log(np.clip(depth, 1.0, 32.0)) / np.log(32.0)            # streetview_encoder.py:112-116
        bi = t * (S - 1)                                                 # (0.5 + t (S-1)) - 0.5
        b0 = int(np.floor(bi)); b1 = min(b0 + 1, S - 1); wb1 = bi - b0
        samp = sum(w * fimg[r, c] for (r, c), w in zip(taps, weights))
        feat, score = samp[:D], (1 - wb1) * samp[D + b0] + wb1 * samp[D + b1]
        dfeat, dscore = rng.standard_normal(D), float(rng.standard_normal())
        ((feat * torch.tensor(dfeat)).sum() + score * dscore).backward()
        g = bf.lift_gather_backward((Hf, Wf, D + S), taps, weights, (b0, b1), wb1, dfeat, dscore, D)
        assert np.abs(g - fimg.grad.numpy()).max() < 1e-12
