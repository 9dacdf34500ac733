"""torch-CPU emulation of the C-ABI operators the training step of the 'resnet_stage' decoder launches, with the
semantics `include/snapb200.h` documents (segment GEMM with row remap / residual / bias / ReLU / row mask / fused GroupNorm
statistics, GroupNorm apply / backward, split-K weight gradients, ...).  TEST INFRASTRUCTURE ONLY: it lets the CPU suite
run the product's launch plan (`snap_b200/semantic_train.py`: which buffer, which offset, which operand layout, in which
order) against torch autograd of the oracle without a GPU; the kernels themselves are checked on the GPU."""
import contextlib

import numpy as np
import torch

from snap_b200 import image_encoder, ops

BF = torch.bfloat16


def _rd(t):
    return t.to(BF).float()


def gemm(a, b, out, *, m_rows=None, n=None, seg_off=(0,), seg_k=None, a_col0=0, residual=None, bias=None, row_mask=None,
         relu=False, remap=None, bn=0, gn_acc=None, gn_acc_relu=None, gn_rows_per_img=0):
    num_seg = len(seg_off)
    seg_k = seg_k if seg_k is not None else b.shape[1] // num_seg
    m_rows = m_rows if m_rows is not None else a.shape[0]
    n = n if n is not None else b.shape[0]
    A, B = a.float(), b.float()
    acc = torch.zeros((m_rows, n), dtype=torch.float32)
    idx0 = torch.arange(m_rows)
    for s, off in enumerate(seg_off):
        idx = idx0 + int(off)
        ok = (idx >= 0) & (idx < A.shape[0])            # TMA zero fill outside the A matrix
        rows = torch.zeros((m_rows, seg_k), dtype=torch.float32)
        rows[ok] = A[idx[ok], a_col0:a_col0 + seg_k]
        acc += rows @ B[:n, s * seg_k:(s + 1) * seg_k].T
    y = _rd(acc)
    if remap is not None:                                # A rows are (img, r, c) of an [R, C] layout -> dense (Ho, Wo) rows
        R, Cc, r0, c0, Ho, Wo = remap
        img, rem = idx0 // (R * Cc), idx0 % (R * Cc)
        r, c = rem // Cc - r0, rem % Cc - c0
        keep = (r >= 0) & (r < Ho) & (c >= 0) & (c < Wo)
        orow = (img * Ho + r) * Wo + c
        y, orow = y[keep], orow[keep]
    else:
        orow = idx0
    if bias is not None:
        y = _rd(y + bias.float()[:n])
    if relu:
        y = torch.relu(y)
    if residual is not None:
        y = _rd(y + residual.float()[orow, :n])
    if row_mask is not None:
        y = y * (row_mask[orow] != 0).float()[:, None]
    y = y.to(out.dtype)
    out[orow, :n] = y
    if gn_acc is not None:
        v = y.double()
        img = orow // gn_rows_per_img
        g = v.reshape(len(orow), 32, n // 32)
        for i in img.unique():
            sel = g[img == i]
            gn_acc[0, i, :, 0] += sel.sum(dim=(0, 2))
            gn_acc[0, i, :, 1] += (sel * sel).sum(dim=(0, 2))
    return out


def gn_stats(x, n, hw, Cc, pre_relu, acc):
    v = x[: n * hw, :Cc].double().reshape(n, hw, 32, Cc // 32)
    if pre_relu:
        v = torch.relu(v)
    acc[0, :, :, 0] += v.sum(dim=(1, 3))
    acc[0, :, :, 1] += (v * v).sum(dim=(1, 3))


def _stats(acc, n, hw, Cc):
    s = acc.sum(dim=0)                                   # replicas
    cnt = hw * (Cc // 32)
    mu = s[..., 0] / cnt
    var = torch.clamp(s[..., 1] / cnt - mu * mu, min=0)
    return mu.float(), (1.0 / torch.sqrt(var + 1e-5)).float()     # [n, 32]


def _xhat(x, n, hw, Cc, acc, pre_relu=False):
    mu, rstd = _stats(acc, n, hw, Cc)
    xg = x[: n * hw, :Cc].float().reshape(n, hw, 32, Cc // 32)
    if pre_relu:
        xg = torch.relu(xg)
    return ((xg - mu[:, None, :, None]) * rstd[:, None, :, None]).reshape(n, hw, Cc), rstd


def _gn_forward(xh, scale, bias, post_relu):
    y = _rd(_rd(_rd(xh) * _rd(scale.float())) + _rd(bias.float()))
    return torch.relu(y) if post_relu else y


def _padded_rows(n, H, W):
    i = torch.arange(n)[:, None, None]
    h = torch.arange(H)[None, :, None]
    w = torch.arange(W)[None, None, :]
    return ((i * (H + 2) + h + 1) * (W + 2) + w + 1).reshape(-1)


def _phase_rows(n, H, W):
    """rows of gn_apply's LAYOUT_PHASE: [2, 2, n, H/2+1, W/2+1] planes of the zero-bordered tensor."""
    Hq, Wq = H // 2 + 1, W // 2 + 1
    i = torch.arange(n)[:, None, None]
    hp = torch.arange(H)[None, :, None] + 1
    wp = torch.arange(W)[None, None, :] + 1
    plane = ((hp & 1) * 2 + (wp & 1)) * (n * Hq * Wq)
    return (plane + (i * Hq + (hp >> 1)) * Wq + (wp >> 1)).reshape(-1)


def gn_apply(x, n, H, W, Cc, acc, scale, bias, pre_relu, post_relu, layout, out, out_sub=None):
    xh, _ = _xhat(x, n, H * W, Cc, acc, pre_relu)
    y = _gn_forward(xh, scale, bias, post_relu).reshape(n * H * W, Cc).to(out.dtype)
    if layout == ops.LAYOUT_DENSE:
        out[: n * H * W, :Cc] = y
    elif layout == ops.LAYOUT_PADDED:
        out[_padded_rows(n, H, W), :Cc] = y
    else:
        out[_phase_rows(n, H, W), :Cc] = y
    if out_sub is not None:                              # even-pixel subsample, dense [n, H/2, W/2, C]
        out_sub[: n * (H // 2) * (W // 2), :Cc] = y.reshape(n, H, W, Cc)[:, ::2, ::2].reshape(-1, Cc)


def gn_backward(x, dy, n, H, W, Cc, acc, scale, bias, accb, dx, dscale, dbias, *, post_relu=True, padded_out=False,
                add=None, pre_relu=False, out_layout=None, dy_phase=False, dy_sub=None):
    hw = H * W
    xh, rstd = _xhat(x, n, hw, Cc, acc, pre_relu)
    if dy_phase:
        d = dy.reshape(-1, Cc)[_phase_rows(n, H, W)].float().reshape(n, hw, Cc)
    else:
        d = dy[: n * hw, :Cc].float().reshape(n, hw, Cc)
    if dy_sub is not None:
        d = d.reshape(n, H, W, Cc).clone()
        d[:, ::2, ::2] = _rd(d[:, ::2, ::2] + dy_sub[: n * (H // 2) * (W // 2), :Cc].float().reshape(n, H // 2, W // 2, Cc))
        d = d.reshape(n, hw, Cc)
    if post_relu:
        d = d * (_gn_forward(xh, scale, bias, False) > 0).float()
    sc = _rd(scale.float())[:Cc]
    sd, sx = d.double().sum(1), (d * xh).double().sum(1)                              # [n, Cc]
    dbias[:Cc] = sd.sum(0).float()
    dscale[:Cc] = sx.sum(0).float()
    m = hw * (Cc // 32)
    s1 = (sd * sc.double()).reshape(n, 32, Cc // 32).sum(-1) / m
    s2 = (sx * sc.double()).reshape(n, 32, Cc // 32).sum(-1) / m
    rep = lambda t: t.float().repeat_interleave(Cc // 32, dim=1)[:, None, :]           # [n, 1, Cc]
    r = rep(rstd) * (d * sc - rep(s1) - xh * rep(s2))
    if pre_relu:
        r = r * (x[: n * hw, :Cc].float().reshape(n, hw, Cc) > 0)
    if add is not None:
        r = r + add[: n * hw, :Cc].float().reshape(n, hw, Cc)
    r = r.reshape(n * hw, Cc).to(dx.dtype)
    if out_layout is None:
        out_layout = 1 if padded_out else 0
    if out_layout == 1:
        dx.view(-1, Cc)[_padded_rows(n, H, W)] = r
    elif out_layout == 2:                                # bottom / right extended [n, H+1, W+1]
        i, h, w = torch.arange(n)[:, None, None], torch.arange(H)[None, :, None], torch.arange(W)[None, None, :]
        dx.view(-1, Cc)[((i * (H + 1) + h) * (W + 1) + w).reshape(-1)] = r
    else:
        dx.view(-1, Cc)[: n * hw] = r


def dense_wgrad(x, dy, M, K, N, dW, db, workspace=None):
    assert M % 16 == 0 and x.shape[0] >= M and dy.shape[0] >= M
    dW[:] = (x[:M, :K].double().T @ dy[:M, :N].double()).float()
    if db is not None:
        db[:N] = dy[:M, :N].double().sum(0).float()


def cast_pad_bf16(src, dst):
    dst.zero_()
    dst[:, : src.shape[1]] = src.to(BF)


def relu_bwd(h, dx, elems):
    m = (h.reshape(-1)[:elems].float() > 0)
    flat = dx.reshape(-1)
    flat[:elems] = torch.where(m, flat[:elems], torch.zeros((), dtype=dx.dtype))


def wt_segments(b_fwd, cout, cin, taps, out):
    for t in range(taps):
        out[:cin, t * cout:(t + 1) * cout] = b_fwd[:cout, t * cin:(t + 1) * cin].T


def stdconv_backward(w, dws, dw):
    w64, d64 = w.double(), dws.double()
    mu = w64.mean(0, keepdim=True)
    rstd = 1.0 / torch.sqrt(((w64 - mu) ** 2).mean(0, keepdim=True) + 1e-10)
    ws = (w64 - mu) * rstd
    dw[:] = (rstd * (d64 - d64.mean(0, keepdim=True) - ws * (d64 * ws).mean(0, keepdim=True))).float()


def _bank_run(self):
    """`_WeightBank.run`: StdConv standardisation (resnet.py:34-41,73-79) + relayout to the bf16 [Cout, K] B operands."""
    off = 0
    for i, (_, k, cout, ldb, std) in enumerate(self.entries):
        w = self.master[off: off + k * cout].view(k, cout).float()
        off += k * cout
        if std:
            w = w - w.mean(0, keepdim=True)
            w = w / torch.sqrt((w * w).mean(0, keepdim=True) + 1e-10)
        self.b_mats[i].zero_()
        self.b_mats[i][:cout, :k] = w.T.to(BF)


def vertical_max_backward(vol, valid, dplane, cells, Z, Cc, dvol):
    v = vol.reshape(-1)[: cells * Z * Cc].float().reshape(cells, Z, Cc)
    m = (valid.reshape(-1)[: cells * Z] != 0).reshape(cells, Z, 1)
    mv = torch.where(m, v, torch.full_like(v, -float("inf")))
    tie = (mv == mv.amax(1, keepdim=True)) & m
    cnt = tie.sum(1, keepdim=True).clamp(min=1)
    g = dplane.reshape(-1)[: cells * Cc].float().reshape(cells, 1, Cc) / cnt
    dvol.reshape(-1)[: cells * Z * Cc] = torch.where(tie, g.expand(cells, Z, Cc), torch.zeros(())).to(dvol.dtype).reshape(-1)


def make_lift_emulation(p2d, vis, depth, more=None):
    """`ops.lift_gather_pool` / `ops.lift_gather_pool_backward` for fixed scenes whose projection (p2d, vis, depth:
    NumPy oracle) is captured here -- one scene, or several told apart by their voxel count (`more`: {N: (p2d, vis,
    depth)}); the float path is tests/lift_torch_ref.py (forward checked against the NumPy oracle, autograd against the
    closed forms of the CUDA kernel, tests/test_lift_backward_ref_cpu.py)."""
    from lift_torch_ref import gather_pool_stats
    scenes = dict(more or {})
    scenes[len(p2d)] = (p2d, vis, depth)

    def lift_gather_pool(lp, views, fimg, xs, ys, zs, stats, valid, dbg_vis=None, dbg_taps=None):
        N = lp.X * lp.Y * lp.Z
        sp2d, svis, sdepth = scenes[N]
        st = gather_pool_stats(fimg.float().reshape(lp.V, lp.Hf, lp.Wf, lp.CF), sp2d, svis, sdepth, lp.D, rd=_rd)
        stats[:N].zero_()
        stats[:N, : st.shape[1]] = st.to(stats.dtype)
        valid[:N] = torch.from_numpy(svis.any(-1).astype(np.uint8))

    def lift_gather_pool_backward(lp, views, fimg, xs, ys, zs, dstats, gimg):
        N = lp.X * lp.Y * lp.Z
        sp2d, svis, sdepth = scenes[N]
        f = fimg.float().reshape(lp.V, lp.Hf, lp.Wf, lp.CF).clone().requires_grad_(True)
        st = gather_pool_stats(f, sp2d, svis, sdepth, lp.D, rd=_rd)
        (st * dstats[:N, : st.shape[1]].float()).sum().backward()
        gimg += f.grad.reshape(gimg.shape)

    return {"lift_gather_pool": lift_gather_pool, "lift_gather_pool_backward": lift_gather_pool_backward}


@contextlib.contextmanager
def emulated_ops(extra=None):
    """Replace the product's operator wrappers (and the weight bank's device pass) by the emulation above."""
    names = ("gemm", "gn_stats", "gn_apply", "gn_backward", "dense_wgrad", "cast_pad_bf16", "relu_bwd", "wt_segments",
             "stdconv_backward", "vertical_max_backward", "match_head_backward", "fuse_max_backward",
             "loc_nll_backward", "loc_pose_scoring_backward", "sem_labels", "sem_loss", "sem_loss_grad", "adam_step",
             "upsample2x", "upsample2x_backward", "root_im2col", "maxpool3x3s2", "maxpool3x3s2_backward")
    table = {n: globals()[n] for n in names}
    table.update(extra or {})
    names = tuple(table)
    saved = {n: getattr(ops, n) for n in names}
    saved_run = image_encoder._WeightBank.run
    try:
        for n in names:
            setattr(ops, n, table[n])
        image_encoder._WeightBank.run = _bank_run
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        image_encoder._WeightBank.run = saved_run


def match_head_backward(plane, valid, cells, Cc, kernel, bias, dout, dy):
    x = plane.reshape(-1)[: cells * Cc].float().reshape(cells, Cc)
    y = _rd(_rd(x @ kernel.float()) + bias.float())
    n = y.norm(dim=-1, keepdim=True)
    z = y / n.clamp(min=1e-30)
    dz = dout.reshape(-1)[: cells * 32].float().reshape(cells, 32)
    g = (dz - z * (z * dz).sum(-1, keepdim=True)) / n.clamp(min=1e-30)
    ok = (valid.reshape(-1)[:cells] != 0)[:, None] & (n >= 1e-5)
    dy.reshape(-1)[: cells * 32] = torch.where(ok, g, torch.zeros(())).to(dy.dtype).reshape(-1)


def fuse_max_backward(a, va, b, vb, dout, cells, Cc, da, db):
    fa, fb = (t.reshape(-1)[: cells * Cc].float().reshape(cells, Cc) for t in (a, b))
    g = dout.reshape(-1)[: cells * Cc].float().reshape(cells, Cc)
    oa = (va.reshape(-1)[:cells] != 0)[:, None]
    ob = torch.ones_like(oa) if vb is None else (vb.reshape(-1)[:cells] != 0)[:, None]
    wa = torch.where(oa & ob, (fa > fb).float() + 0.5 * (fa == fb).float(), oa.float().expand_as(fa))
    wb = torch.where(oa & ob, (fb > fa).float() + 0.5 * (fa == fb).float(), (ob & ~oa).float().expand_as(fa))
    da.reshape(-1)[: cells * Cc] = (wa * g).to(da.dtype).reshape(-1)
    db.reshape(-1)[: cells * Cc] = (wb * g).to(db.dtype).reshape(-1)


def loc_nll_backward(scores, remove, dr_samples, dt_samples, dscores, dtemperature=None):
    B, P1 = scores.shape
    s = scores.clone().requires_grad_(True)
    rem = torch.zeros((B, P1), dtype=torch.bool)
    if remove is not None:
        rem = (dr_samples < remove[0]) & (dt_samples < remove[1])
        rem[:, 0] = False
    sc = torch.where(rem, torch.full((), -float("inf")), s)
    (torch.logsumexp(sc, 1) - sc[:, 0]).mean().backward()
    dscores[:] = s.grad
    if dtemperature is not None:
        dtemperature[:] = (s.grad * scores).sum(1)


def loc_pose_scoring_backward(sim, point_scale, i_xy, valid_j, poses, dscores, H, W, cell_size, mask_out_of_bounds,
                              relu_mask, dsim):
    from loc_torch_ref import pose_scores, pose_uv
    B, N = point_scale.shape
    for b in range(B):
        s = sim[b].float().reshape(N, H, W).clone().requires_grad_(True)
        uv = pose_uv(poses[b].numpy(), (i_xy[b] if i_xy.dim() == 3 else i_xy).numpy(), cell_size)
        sp = s * point_scale[b][:, None, None]
        sc = pose_scores(sp, uv, np.ones(N, bool), None if valid_j is None else valid_j[b].numpy().astype(bool).reshape(H, W),
                         mask_out_of_bounds)
        (sc * dscores[b]).sum().backward()
        g = s.grad.reshape(N, H * W)
        if relu_mask:
            g = g * (sim[b].float() > 0)
        dsim[b, :N] = g.to(dsim.dtype)


# ---- semantic losses / optimizer (GPU-verified kernels; emulated here so that a whole train_step runs on the CPU) --------
def sem_labels(sel_area, sel_excl, sel_indep, num_gt, masks, bev_valid, labels_area, valid_area, labels_excl, masks_indep):
    m = masks.reshape(-1, num_gt) != 0

    def pick(sel):
        return torch.stack([m[:, list(s)].any(1) for s in sel], 1)                # 'line' absorbs its siblings (:259-266)
    a = pick(sel_area)
    labels_area.reshape(-1)[:] = a.int().argmax(1).int()
    valid_area.reshape(-1)[:] = a.any(1).to(torch.uint8)
    if labels_excl is not None:
        e = pick(sel_excl)
        labels_excl.reshape(-1)[:] = torch.where(e.any(1), e.int().argmax(1), torch.tensor(len(sel_excl))).int()   # void
        masks_indep.reshape(m.shape[0], -1)[:] = m[:, list(sel_indep)].to(torch.uint8)


def _total_loss(logits, labels_area, valid_area, labels_excl, masks_indep, valid, Ka, Ke, Ki, weights):
    from oracle import semantic_net as osn
    B, cells, _ = logits.shape
    r = lambda t: t.reshape(B, 1, cells, *t.shape[2:]).numpy() if t is not None else None
    w = [None if x is None else x.numpy() for x in (weights or [None] * 4)]
    return osn.total_loss_torch(logits[..., : Ka + Ke + Ki].reshape(B, 1, cells, -1), r(labels_area), r(valid_area) != 0,
                                r(labels_excl), r(masks_indep), r(valid) != 0, Ka, Ke, *w)


def sem_loss(logits, labels_area, valid_area, labels_excl, masks_indep, valid, Ka, Ke, Ki, weights, out):
    out.zero_()
    out[:, 3] = _total_loss(logits, labels_area, valid_area, labels_excl, masks_indep, valid, Ka, Ke, Ki, weights)[1]


def sem_loss_grad(logits, labels_area, valid_area, labels_excl, masks_indep, valid, Ka, Ke, Ki, weights, counts, dlogits):
    lg = logits.clone().requires_grad_(True)
    _total_loss(lg, labels_area, valid_area, labels_excl, masks_indep, valid, Ka, Ke, Ki, weights)[0].backward()
    B, cells, ld = logits.shape
    dlogits.zero_()
    dlogits[: B * cells, :ld] = lg.grad.reshape(B * cells, ld).to(dlogits.dtype)


def adam_step(p, m, v, g, lr, step, b1=0.9, b2=0.999, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr * (m / (1 - b1 ** step)) / (torch.sqrt(v / (1 - b2 ** step)) + eps))


def upsample2x(x, n, h, w, Cc, y):
    v = x.reshape(-1)[: n * h * w * Cc].float().reshape(n, h, w, Cc).permute(0, 3, 1, 2)
    up = torch.nn.functional.interpolate(v, scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    y.reshape(-1)[: 4 * n * h * w * Cc] = up.reshape(-1).to(y.dtype)


def upsample2x_backward(dy, n, h, w, Cc, dx):
    v = torch.zeros((n, Cc, h, w), requires_grad=True)
    up = torch.nn.functional.interpolate(v, scale_factor=2, mode="bilinear", align_corners=False)
    g = dy.reshape(-1)[: 4 * n * h * w * Cc].float().reshape(n, 2 * h, 2 * w, Cc).permute(0, 3, 1, 2)
    (up * g).sum().backward()
    dx.reshape(-1)[: n * h * w * Cc] = v.grad.permute(0, 2, 3, 1).reshape(-1).to(dx.dtype)


def root_im2col(images, Hp, Wp, KH, KW, stride, pad, out):
    """bf16 [n*Ho*Wo, Kp]: windows of 2 * bf16(image) - 1 (resnet.py:199; -1 in pad_to_multiple's zero padding, 0 in the
    convolution's own padding), K ordered (kh, kw, c) like the Flax HWIO kernel."""
    n, H, W, _ = images.shape
    x = torch.full((n, Hp, Wp, 3), -1.0)
    x[:, :H, :W] = _rd(_rd(images.float()) * 2 - 1)
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pad, pad, pad, pad))
    Ho, Wo = (Hp + 2 * pad - KH) // stride + 1, (Wp + 2 * pad - KW) // stride + 1
    cols = torch.nn.functional.unfold(xp, (KH, KW), stride=stride)              # [n, 3*KH*KW (c, kh, kw), Ho*Wo]
    cols = cols.reshape(n, 3, KH, KW, Ho * Wo).permute(0, 4, 2, 3, 1).reshape(n * Ho * Wo, KH * KW * 3)
    out[: n * Ho * Wo].zero_()
    out[: n * Ho * Wo, : KH * KW * 3] = cols.to(out.dtype)


def maxpool3x3s2(x, n, H, W, Cc, y):
    v = x.reshape(-1)[: n * H * W * Cc].float().reshape(n, H, W, Cc).permute(0, 3, 1, 2)
    o = torch.nn.functional.max_pool2d(v, 3, stride=2, padding=1).permute(0, 2, 3, 1)
    y.reshape(-1)[: o.numel()] = o.reshape(-1).to(y.dtype)


def maxpool3x3s2_backward(x, dy, n, H, W, Cc, dx):
    v = x.reshape(-1)[: n * H * W * Cc].float().reshape(n, H, W, Cc).permute(0, 3, 1, 2).clone().requires_grad_(True)
    o = torch.nn.functional.max_pool2d(v, 3, stride=2, padding=1)
    g = dy.reshape(-1)[: o.numel()].float().reshape(n, o.shape[2], o.shape[3], Cc).permute(0, 3, 1, 2)
    (o * g).sum().backward()
    dx.reshape(-1)[: n * H * W * Cc] = v.grad.permute(0, 2, 3, 1).reshape(-1).to(dx.dtype)
