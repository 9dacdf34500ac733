"""Image-encoder kernels and the full ImageEncoder vs the oracle (oracle/resnet.py, oracle/image_encoder.py)."""
import numpy as np
import pytest
import torch

from util import F, assert_close_bf16, bf16_np, rd_bf16, rel_l2, record_parity

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=F))


def test_std_weights_vs_oracle():
    from oracle import resnet as ores
    from snap_b200.image_encoder import _WeightBank
    rng = np.random.default_rng(0)
    shapes = [(7, 7, 3, 64), (1, 1, 64, 256), (3, 3, 128, 128), (1, 1, 2048, 128), (257, 256)]
    stds = [True, True, True, False, False]
    bank = _WeightBank(torch.device("cuda"))
    ks = [bank.add((rng.standard_normal(s) * 0.1 + 0.02).astype(F), st, 32) for s, st in zip(shapes, stds)]
    bank.finalize()
    bank.run()
    torch.cuda.synchronize()
    for i, (s, st) in enumerate(zip(shapes, stds)):
        w = bank.entries[i][0]  # [K, Cout]
        ref = ores.std_kernel(_t(w).reshape(*s)).reshape(-1, s[-1]).numpy() if st else w
        out = bank.b_mats[i].float().cpu().numpy()
        K = w.shape[0]
        assert_close_bf16(out[: s[-1], :K], ref.T, f"std_weights{s}", atol_scale=1e-4)
        assert not out[:, K:].any()


def test_gn_maxpool_upsample_im2col_vs_oracle():
    from oracle import resnet as ores
    from snap_b200 import ops
    import torch.nn.functional as Fnn
    rng = np.random.default_rng(1)
    dev = "cuda"
    # GroupNorm stats + apply in all three layouts
    for (n, h, w, c) in [(2, 12, 20, 64), (1, 6, 10, 2048), (3, 8, 8, 256)]:
        x = bf16_np(rng.standard_normal((n, h, w, c)) * 2 + 0.5)
        scale, bias = (1 + 0.3 * rng.standard_normal(c)).astype(F), (0.2 * rng.standard_normal(c)).astype(F)
        ref = torch.relu(ores.group_norm(_t(x), _t(scale), _t(bias))).numpy()
        xd = _t(x).to(torch.bfloat16).to(dev)
        stats = torch.zeros((8, n, 32, 2), dtype=torch.float64, device=dev)
        ops.gn_stats(xd, n, h * w, c, False, stats)
        dense = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=dev)
        sub = torch.zeros((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=dev)
        ops.gn_apply(xd, n, h, w, c, stats, _t(scale).to(dev), _t(bias).to(dev), False, True, ops.LAYOUT_DENSE, dense, sub)
        padded = torch.zeros((n, h + 2, w + 2, c), dtype=torch.bfloat16, device=dev)
        ops.gn_apply(xd, n, h, w, c, stats, _t(scale).to(dev), _t(bias).to(dev), False, True, ops.LAYOUT_PADDED, padded)
        phase = torch.zeros((2, 2, n, h // 2 + 1, w // 2 + 1, c), dtype=torch.bfloat16, device=dev)
        ops.gn_apply(xd, n, h, w, c, stats, _t(scale).to(dev), _t(bias).to(dev), False, True, ops.LAYOUT_PHASE, phase)
        torch.cuda.synchronize()
        assert_close_bf16(dense.float().cpu().numpy(), ref, f"gn dense {c}")
        assert_close_bf16(sub.float().cpu().numpy(), ref[:, ::2, ::2], f"gn sub {c}")
        refp = np.zeros((n, h + 2, w + 2, c), F)
        refp[:, 1:-1, 1:-1] = ref
        assert np.array_equal(padded.float().cpu().numpy()[:, 1:-1, 1:-1], dense.float().cpu().numpy())
        assert not padded.float().cpu().numpy()[:, 0].any() and not padded.float().cpu().numpy()[:, :, -1].any()
        ph = phase.float().cpu().numpy()
        dn = np.zeros((n, h + 2, w + 2, c), F)
        dn[:, 1:-1, 1:-1] = dense.float().cpu().numpy()
        for a in range(2):
            for b in range(2):
                assert np.array_equal(ph[a, b], dn[:, a::2, b::2]), (a, b)
    # FPN-style GN: relu first, no relu after
    n, h, w, c = 1, 8, 8, 512
    x = bf16_np(rng.standard_normal((n, h, w, c)))
    scale, bias = (1 + 0.3 * rng.standard_normal(c)).astype(F), (0.2 * rng.standard_normal(c)).astype(F)
    ref = ores.group_norm(torch.relu(_t(x)), _t(scale), _t(bias)).numpy()
    xd = _t(x).to(torch.bfloat16).to(dev)
    stats = torch.zeros((8, n, 32, 2), dtype=torch.float64, device=dev)
    out = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=dev)
    ops.gn_stats(xd, n, h * w, c, True, stats)
    ops.gn_apply(xd, n, h, w, c, stats, _t(scale).to(dev), _t(bias).to(dev), True, False, ops.LAYOUT_DENSE, out)
    torch.cuda.synchronize()
    assert_close_bf16(out.float().cpu().numpy(), ref, "gn fpn")
    # max pool
    x = bf16_np(rng.standard_normal((2, 16, 24, 64)))
    ref = Fnn.max_pool2d(_t(x).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).numpy()
    y = torch.zeros((2, 8, 12, 64), dtype=torch.bfloat16, device=dev)
    ops.maxpool3x3s2(_t(x).to(torch.bfloat16).to(dev), 2, 16, 24, 64, y)
    torch.cuda.synchronize()
    assert np.array_equal(y.float().cpu().numpy(), ref)
    # x2 bilinear
    x = bf16_np(rng.standard_normal((2, 5, 7, 128)))
    ref = Fnn.interpolate(_t(x).permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1).numpy()
    y = torch.zeros((2, 10, 14, 128), dtype=torch.bfloat16, device=dev)
    ops.upsample2x(_t(x).to(torch.bfloat16).to(dev), 2, 5, 7, 128, y)
    torch.cuda.synchronize()
    assert_close_bf16(y.float().cpu().numpy(), ref, "upsample2x")
    # root im2col (+ pad semantics: padded region is -1 after 2x-1, conv padding is 0)
    img = rng.random((2, 10, 14, 3), dtype=F)
    Hp, Wp = 32, 32
    a = torch.zeros((2 * 16 * 16, 160), dtype=torch.bfloat16, device=dev)
    ops.root_im2col(_t(img).to(dev), Hp, Wp, 7, 7, 2, 3, a)
    torch.cuda.synchronize()
    xb = torch.full((2, Hp, Wp, 3), -1.0)
    xb[:, :10, :14] = rd_bf16(rd_bf16(_t(img)) * 2 - 1)
    cols = Fnn.unfold(xb.permute(0, 3, 1, 2), 7, padding=3, stride=2)  # [n, 3*49, L] channel-major
    cols = cols.reshape(2, 3, 49, -1).permute(0, 3, 2, 1).reshape(2 * 256, 147).numpy()
    out = a.float().cpu().numpy()
    assert np.array_equal(out[:, :147], cols) and not out[:, 147:].any()


@pytest.mark.parametrize("skip_root,hw,fused_gn", [(False, (40, 72), False), (True, (24, 24), False),
                                                   (False, (40, 72), True), (True, (24, 24), True),
                                                   (False, (40, 72), "1x1"), (True, (24, 24), "1x1"),
                                                   (False, (40, 72), "auto")])
def test_image_encoder_units_teacher_forced(skip_root, hw, fused_gn):
    """Every residual unit, the root block and the FPN, each fed the ORACLE's (bf16-mode) input of that
    block, vs the oracle's output of that block.  Both sides round at the same points, so the only
    differences are fp32 summation order and the rare bf16 rounding flips they cause:
    relative L2 <= 4e-3 per block (bf16 eps = 7.8e-3)."""
    from oracle import image_encoder as oie
    from snap_b200 import configs, image_encoder, params
    rng = np.random.default_rng(2)
    cfg = configs.aerial_encoder() if skip_root else configs.image_encoder()
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_image_encoder(rng, cfg)))
    img = rng.random((2, *hw, 3), dtype=F)
    tt = lambda tree: {k: (tt(v) if isinstance(v, dict) else _t(v)) for k, v in tree.items()}
    trace = []
    ref_bf, strides = oie.image_encoder(_t(img), tt(p), skip_root, rd_bf16, trace)
    enc = image_encoder.ImageEncoder(cfg, fused_gn=fused_gn)  # fused_gn: GroupNorm fused into the conv's A path
    plan = enc.plan(p, 2, *hw, torch.device("cuda"))
    plan.bank.run()
    plan.gn_acc_all.zero_()
    x0 = plan.run_root(_t(img).cuda())
    torch.cuda.synchronize()
    rows0 = trace[0][1].numel() // trace[0][1].shape[-1]
    e = rel_l2(x0[:rows0].float().cpu().numpy(), trace[0][1].reshape(rows0, -1).numpy())
    print(f"root: rel_l2 {e:.5f}")
    record_parity("encoder (teacher-forced)", f"root block, fused_gn={fused_gn}", e, 1e-4)
    assert e < 1e-4            # measured 0 (a single bf16 rounding of an fp32-accumulated K = 147 dot product)
    worst = 0.0
    for i, u in enumerate(plan.units):
        xin, yout = trace[i + 1]
        rows_in, rows_out = xin.numel() // xin.shape[-1], yout.numel() // yout.shape[-1]
        xd = torch.zeros((max(128, -(-rows_in // 128) * 128), xin.shape[-1]), dtype=torch.bfloat16, device="cuda")
        xd[:rows_in] = xin.reshape(rows_in, -1).to(torch.bfloat16).cuda()
        out = plan.run_unit(u, xd, forced_input=True)
        torch.cuda.synchronize()
        e = rel_l2(out[:rows_out].float().cpu().numpy(), yout.reshape(rows_out, -1).numpy())
        worst = max(worst, e)
        print(f"unit {i} (cin {u['cin']} stride {u['stride']} {u['h']}x{u['w']}): rel_l2 {e:.5f}")
        # measured <= 1.9e-3 over the 16 units (bf16 flips of the three GroupNorm inputs); tolerance = 1.5 x that
        assert e < 3e-3, f"unit {i}"
    record_parity("encoder (teacher-forced)", f"worst of the 16 bottleneck units, fused_gn={fused_gn}", worst, 3e-3)
    # FPN, teacher-forced with the oracle's stage outputs
    ends = np.cumsum(plan.blocks)
    for (buf, h, w, c), end in zip(plan.stage_out, ends):
        y = trace[end][1]
        buf[: y.numel() // c] = y.reshape(-1, c).to(torch.bfloat16).cuda()
    outs = plan.run_fpn(forced_input=True)
    torch.cuda.synchronize()
    for lvl, (o, (hh, ww), rb) in enumerate(zip(outs, plan.cropped_shapes(), ref_bf)):
        e = rel_l2(o[:, :hh, :ww].float().cpu().numpy(), rb.numpy())
        print(f"fpn level {lvl}: rel_l2 {e:.5f}")
        record_parity("encoder (teacher-forced)", f"FPN level {lvl}, fused_gn={fused_gn}", e, 3e-4)
        assert e < 3e-4          # measured <= 1e-4


@pytest.mark.parametrize("skip_root,hw", [(False, (40, 72)), (True, (24, 24))])
def test_image_encoder_end_to_end(skip_root, hw):
    """Free-running ImageEncoder vs the oracle.  A 50-layer bf16 network with random weights amplifies
    every flipped rounding through ~100 GroupNorms (the oracle's own bf16 mode is 10-20 % away from its
    fp32 mode on these tiny maps), so the end-to-end bound is relative: the CUDA result must be no farther
    from the fp32 oracle than 1.5x the distance of the reference's own bf16 mode, per pyramid level."""
    from oracle import image_encoder as oie
    from snap_b200 import configs, image_encoder, params
    rng = np.random.default_rng(2)
    cfg = configs.aerial_encoder() if skip_root else configs.image_encoder()
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_image_encoder(rng, cfg)))
    img = rng.random((2, *hw, 3), dtype=F)
    tt = lambda tree: {k: (tt(v) if isinstance(v, dict) else _t(v)) for k, v in tree.items()}
    ref_bf, strides = oie.image_encoder(_t(img), tt(p), skip_root, rd_bf16)
    ref_32, _ = oie.image_encoder(_t(img), tt(p), skip_root)
    enc = image_encoder.ImageEncoder(cfg)
    pyr = enc.apply({"params": p}, _t(img).cuda())
    torch.cuda.synchronize()
    assert len(pyr.features) == len(ref_bf)
    for lvl, (o, rb, r32, s, s_ref) in enumerate(zip(pyr.features, ref_bf, ref_32, pyr.strides, strides)):
        assert tuple(o.shape) == tuple(rb.shape), (lvl, o.shape, rb.shape)
        assert tuple(s) == tuple(s_ref)
        got = o.float().cpu().numpy()
        e_bf, e_32, e_ref = rel_l2(got, rb.numpy()), rel_l2(got, r32.numpy()), rel_l2(rb.numpy(), r32.numpy())
        print(f"level {lvl}: vs bf16-oracle {e_bf:.4f}, vs fp32-oracle {e_32:.4f}, bf16-oracle vs fp32-oracle {e_ref:.4f}")
        assert e_32 < 1.5 * e_ref + 1e-3


@pytest.mark.parametrize("KH,stride,pad,hw", [(7, 2, 3, (70, 300)), (3, 1, 1, (40, 150))])
def test_implicit_root_conv_vs_im2col_gemm(KH, stride, pad, hw):
    """The sliding-window (4D TMA) root conv equals the explicit im2col + GEMM of the same bf16 operands, including
    the pad_to_multiple (-1) and conv padding (0) rings and several 128-column blocks per output row."""
    from snap_b200 import ops
    rng = np.random.default_rng(21)
    n, (H, W) = 3, hw
    mult = 32 if stride == 2 else 8
    Hp, Wp = H + (mult - H % mult), W + (mult - W % mult)
    img = _t(rng.random((n, H, W, 3), dtype=F)).cuda()
    cout = 64
    K = KH * KH * 3
    ldb = (K + 31) // 32 * 32
    b_std = torch.zeros((cout, ldb), dtype=torch.bfloat16, device="cuda")
    b_std[:, :K] = (torch.randn((cout, K), device="cuda") * 0.1).to(torch.bfloat16)
    cp, Hq, Wq, Ho, Wo = ops.root_packed_geometry(Hp, Wp, KH, KH, stride, pad)
    # reference: explicit im2col + GEMM
    a = torch.zeros(((n * Ho * Wo + 127) // 128 * 128, ldb), dtype=torch.bfloat16, device="cuda")
    ops.root_im2col(img, Hp, Wp, KH, KH, stride, pad, a)
    ref = torch.zeros((a.shape[0], cout), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b_std, ref, m_rows=n * Ho * Wo, seg_k=ldb)
    # implicit
    packed = torch.zeros(n * Hq * Wq * cp + 64, dtype=torch.bfloat16, device="cuda")
    ops.root_pack_image(img, Hp, Wp, pad, cp, Hq, Wq, packed)
    bw = torch.zeros((cout, KH * 32), dtype=torch.bfloat16, device="cuda")
    ops.root_pack_weights(b_std, cout, KH, KH, cp, bw)
    out = torch.full((a.shape[0], cout), 7.0, dtype=torch.bfloat16, device="cuda")
    acc = torch.zeros((ops.GN_REPLICAS, n, 32, 2), dtype=torch.float64, device="cuda")
    ops.root_conv(packed, n, Hq, Wq, cp, KH, stride, Ho, Wo, bw, cout, out, gn_acc=acc)
    torch.cuda.synchronize()
    got, want = out[: n * Ho * Wo].float().cpu().numpy(), ref[: n * Ho * Wo].float().cpu().numpy()
    assert_close_bf16(got, want, "implicit root conv")
    assert (out[n * Ho * Wo:] == 7.0).all()                       # nothing written past the last pixel
    s = acc.sum(0).cpu().numpy()                                   # fused statistics of what was stored
    g = got.reshape(n, Ho * Wo, 32, cout // 32).astype(np.float64)
    assert np.allclose(s[..., 0], g.sum((1, 3)), rtol=1e-6, atol=1e-3)
    assert np.allclose(s[..., 1], (g * g).sum((1, 3)), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("n_img,hw,C,N,res,pre,post,relu_acc", [
    (3, (40, 64), 256, 64, False, False, True, False),    # conv1 of stage 1; 2560 rows per image = 20 whole tiles
    (3, (33, 40), 64, 256, True, False, True, True),      # conv3 + residual + both statistics; tiles straddle images
    (2, (30, 44), 512, 128, False, False, True, False),   # conv1 of stage 2 (K = 8 blocks)
    (5, (7, 9), 128, 512, True, False, True, False),      # images smaller than a tile
    (2, (32, 48), 1024, 128, True, True, False, False),   # FPN skip conv: relu -> GN -> conv + up-sampled level
    (2, (24, 40), 2048, 512, False, False, True, False),  # conv1 of stage 4: deep K loop (one CTA per SM)
    (40, (8, 16), 256, 64, False, False, True, False),    # more than 32 images per call (12+ tiles per step)
    (3, (20, 32), 1024, 256, False, False, True, False),  # conv1 of stage 3: 128 x 256 tiles
])
def test_conv_gn_1x1_equals_apply_then_gemm(n_img, hw, C, N, res, pre, post, relu_acc):
    """The A_TGN1 mode of `snapb200_conv_gn_bf16` (GroupNorm + ReLU applied to the raw tile in shared memory, conv
    epilogue) against the two-launch path it replaces (`gn_apply` -> `gemm`): same rounding chain and the same
    K order, so the stored outputs must be BIT-identical; the statistics of the output agree to double rounding."""
    from snap_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(C + N)
    dev = torch.device("cuda")
    H, W = hw
    rows = n_img * H * W
    x = (torch.randn((rows, C), device=dev, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    b = (torch.randn((N, C), device=dev, generator=g) * C ** -0.5).to(torch.bfloat16)
    r = torch.randn((rows, N), device=dev, generator=g).to(torch.bfloat16) if res else None
    scale = (1 + 0.2 * torch.randn(C, device=dev, generator=g)).to(torch.bfloat16).float()
    bias = (0.2 * torch.randn(C, device=dev, generator=g)).to(torch.bfloat16).float()
    acc = torch.zeros((ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64, device=dev)
    ops.gn_stats(x, n_img, H * W, C, pre, acc)
    outs, stats = [], []
    for fused in (False, True):
        out = torch.full((rows + 128, N), 7.0, dtype=torch.bfloat16, device=dev)
        a1 = torch.zeros((ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64, device=dev)
        a2 = torch.zeros_like(a1)
        if fused:
            ops.conv_gn(x, n_img, H, W, C, acc, scale, bias, b, out, pre_relu=pre, post_relu=post, residual=r,
                        gn_acc=a1, gn_acc_relu=a2 if relu_acc else None)
        else:
            a = torch.zeros((rows + 128, C), dtype=torch.bfloat16, device=dev)
            ops.gn_apply(x, n_img, H, W, C, acc, scale, bias, pre, post, ops.LAYOUT_DENSE, a)
            ops.gemm(a, b, out, m_rows=rows, residual=r, gn_acc=a1, gn_acc_relu=a2 if relu_acc else None,
                     gn_rows_per_img=H * W)
        torch.cuda.synchronize()
        outs.append(out.view(torch.int16).cpu().numpy())
        stats.append((a1.sum(0).cpu().numpy(), a2.sum(0).cpu().numpy()))
    assert np.array_equal(outs[0][:rows], outs[1][:rows])
    assert (outs[1][rows:] == outs[0][rows:]).all()  # rows beyond M stay untouched
    for s0, s1 in zip(stats[0], stats[1]):
        np.testing.assert_allclose(s1, s0, rtol=1e-12, atol=1e-9)
