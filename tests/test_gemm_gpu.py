"""tcgen05 GEMM engine vs a plain fp32 reference of the same op (bf16 inputs, fp32 accumulate).

Tolerance: outputs are fp32 sums of exactly-representable bf16 products, so only the summation
order differs: |err| <= 1e-4 * max|ref| (fp32 output) and one bf16 ulp on top for bf16 outputs.
"""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


def _check(out, ref, bf16_out=False):
    out = out.float().cpu()
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-4 * scale + (2.0 ** -8) * ref.abs() * (1.0 if bf16_out else 0.0)
    bad = (out - ref).abs() > tol + 1e-6
    assert not bad.any(), f"{int(bad.sum())} mismatches, max err {(out - ref).abs().max().item()} scale {scale}"


@pytest.mark.parametrize("M,K,N", [
    (128, 64, 64), (300, 128, 64), (1000, 256, 128), (777, 512, 256), (4096, 1024, 512),
    (129, 128, 160), (640, 64, 16), (500, 96, 64), (500, 160, 64), (900, 288, 256), (260, 288, 128),
    (128 * 320 + 5, 64, 256),
])
def test_gemm_plain_f32(M, K, N):
    from snap_b200 import ops
    a, b = _rand((M, K), 1), _rand((N, K), 2)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a.cuda(), b.cuda(), out)
    torch.cuda.synchronize()
    _check(out, a.float() @ b.float().T)


def test_gemm_epilogue_bias_residual_relu_mask():
    from snap_b200 import ops
    M, K, N = 700, 256, 128
    a, b = _rand((M, K), 3), _rand((N, K), 4)
    res = _rand((M, N), 5)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6))
    mask = (torch.rand(M, generator=torch.Generator().manual_seed(7)) > 0.3).to(torch.uint8)
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    ops.gemm(a.cuda(), b.cuda(), out, residual=res.cuda(), bias=bias.cuda(), row_mask=mask.cuda(), relu=True)
    torch.cuda.synchronize()
    r = lambda t: t.to(torch.bfloat16).float()  # the epilogue rounds like the reference: dot, +bias, +residual
    ref = torch.relu(r(r(a.float() @ b.float().T) + bias) + res.float()) * mask[:, None].float()
    _check(out, ref, bf16_out=True)


@pytest.mark.parametrize("stride", [1, 2])
def test_gemm_conv3x3_segments(stride):
    """3x3 conv = 9 row-shifted K-segments over a zero-bordered (stride 1) or phase-split (stride 2) buffer."""
    from snap_b200 import ops
    n_img, H, W, Cin, Cout = 2, 12, 20, 64, 128
    x = _rand((n_img, H, W, Cin), 8)
    w = _rand((3, 3, Cin, Cout), 9, 0.1)  # HWIO
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(3, 2, 0, 1),
                                     stride=stride, padding=1).permute(0, 2, 3, 1)
    bmat = w.permute(3, 0, 1, 2).reshape(Cout, 9 * Cin).contiguous()
    xp = torch.zeros((n_img, H + 2, W + 2, Cin), dtype=torch.bfloat16)
    xp[:, 1:-1, 1:-1] = x
    if stride == 1:
        Hp, Wp = H + 2, W + 2
        a = xp.reshape(-1, Cin)
        seg_off = [(kh - 1) * Wp + (kw - 1) for kh in range(3) for kw in range(3)]
        remap = (Hp, Wp, 1, 1, H, W)
        m_rows = n_img * Hp * Wp
        Ho, Wo = H, W
    else:
        Ho, Wo = H // 2, W // 2
        Hq, Wq = Ho + 1, Wo + 1
        planes = torch.stack([xp[:, a_::2, b_::2] for a_ in range(2) for b_ in range(2)])  # [4,n,Hq,Wq,C]
        assert planes.shape[2:4] == (Hq, Wq)
        a = planes.reshape(-1, Cin)
        plane_rows = n_img * Hq * Wq
        seg_off = [((kh % 2) * 2 + (kw % 2)) * plane_rows + (kh // 2) * Wq + (kw // 2)
                   for kh in range(3) for kw in range(3)]
        remap = (Hq, Wq, 0, 0, Ho, Wo)
        m_rows = plane_rows
    out = torch.full((n_img * Ho * Wo, Cout), float("nan"), device="cuda")
    ops.gemm(a.contiguous().cuda(), bmat.cuda(), out, m_rows=m_rows, seg_off=seg_off, seg_k=Cin, remap=remap)
    torch.cuda.synchronize()
    _check(out, ref.reshape(-1, Cout))


def test_gemm_rejects_bad_shapes():
    from snap_b200 import _lib, ops
    a, b = _rand((128, 48), 1).cuda(), _rand((64, 48), 2).cuda()
    out = torch.zeros((128, 64), device="cuda")
    with pytest.raises(_lib.SnapB200Error):
        ops.gemm(a, b, out)  # K=48 is not a multiple of 32


@pytest.mark.parametrize("n,cpg_n", [(64, 64), (256, 256), (2048, 2048)])
def test_gemm_epilogue_groupnorm_accumulators(n, cpg_n):
    """Fused GroupNorm statistics: acc[img][g] = (sum, sumsq) of the STORED bf16 output (with residual),
    including M-tiles that straddle an image boundary and rows beyond M."""
    from snap_b200 import ops
    rows_per_img, n_img, K = 200, 3, 128
    M = rows_per_img * n_img
    a, b = _rand((M, K), 21), _rand((n, K), 22, 0.2)
    res = _rand((M, n), 23)
    out = torch.zeros((M, n), dtype=torch.bfloat16, device="cuda")
    acc = torch.zeros((8, n_img, 32, 2), dtype=torch.float64, device="cuda")
    acc_relu = torch.zeros((8, n_img, 32, 2), dtype=torch.float64, device="cuda")
    ops.gemm(a.cuda(), b.cuda(), out, residual=res.cuda(), gn_acc=acc, gn_acc_relu=acc_relu, gn_rows_per_img=rows_per_img)
    torch.cuda.synchronize()
    r = lambda t: t.to(torch.bfloat16).float()
    ref = r(r(a.float() @ b.float().T) + res.float())
    assert torch.equal(out.float().cpu(), ref) or (out.float().cpu() - ref).abs().max() <= 2.0 ** -7 * ref.abs().max()
    o = out.float().cpu().double().reshape(n_img, rows_per_img, 32, n // 32)
    for accd, x in ((acc, o), (acc_relu, o.clamp(min=0))):
        s, q = x.sum(dim=(1, 3)), (x * x).sum(dim=(1, 3))
        assert torch.allclose(accd.cpu().sum(0)[..., 0], s, rtol=1e-5, atol=1e-3)
        assert torch.allclose(accd.cpu().sum(0)[..., 1], q, rtol=1e-5, atol=1e-3)


def test_row_shifted_swizzled_descriptor():
    """The halo form of the 3x3 conv reads its nine taps from ONE shared-memory block by moving the A descriptor's start by
    whole rows (128 B).  Shown here on the device, bit-exactly (integer data): with SWIZZLE_128B the hardware derives the
    XOR phase from the ABSOLUTE shared-memory address, so any start row works with the matrix-base-offset field left 0
    (the block is 1024-byte aligned and was written by TMA with the same address-based pattern); setting the field to
    (address >> 7) & 7 breaks every start row that is not a multiple of 8."""
    from snap_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-4, 5, (256, 64), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randint(-4, 5, (64, 64), device="cuda", generator=g).to(torch.bfloat16)
    shifts = (0, 1, 3, 7, 8, 13, 64, 100, 127, 128)
    ok0, ok1 = [], []
    for shift in shifts:
        ref = a[shift:shift + 128].float() @ b.float().t()
        r0 = ops.selftest_shifted_desc(a, b, shift, 0)
        r1 = ops.selftest_shifted_desc(a, b, shift, 1)
        torch.cuda.synchronize()
        ok0.append(bool(torch.equal(r0, ref)))
        ok1.append(bool(torch.equal(r1, ref)))
    print("shift        :", shifts)
    print("base_offset=0:", ok0)
    print("base_offset  :", ok1)
    assert all(ok0), "row-shifted descriptors (base offset 0) must address the shifted rows"
    assert all(o == (sh % 8 == 0) for o, sh in zip(ok1, shifts)), "the base-offset field is not an address correction"


@pytest.mark.parametrize("n_img,H,W,C,N", [(2, 24, 40, 64, 64), (3, 17, 23, 64, 64), (1, 128, 168, 64, 64),
                                           (2, 16, 20, 128, 128), (1, 30, 44, 128, 64), (2, 64, 84, 128, 128)])
def test_conv3x3_halo_equals_nine_segment_gemm(n_img, H, W, C, N):
    """`snapb200_conv3x3_halo_bf16` (one shared-memory halo block per K chunk, nine row-shifted descriptors, resident
    weights) against the 9-segment form of the GEMM engine on the same zero-bordered input: identical products; for
    C = 64 also the same accumulation order, i.e. BIT-identical outputs; GroupNorm statistics to double rounding."""
    from snap_b200 import ops
    assert ops.conv3x3_halo_supported(C, N, W)
    g = torch.Generator(device="cuda").manual_seed(H * W + C)
    dev = torch.device("cuda")
    hp, wp = H + 2, W + 2
    a = torch.zeros((n_img, hp, wp, C), dtype=torch.bfloat16, device=dev)
    a[:, 1:-1, 1:-1] = torch.randn((n_img, H, W, C), device=dev, generator=g).to(torch.bfloat16)
    rows_p = n_img * hp * wp
    a_flat = torch.zeros((rows_p + 256, C), dtype=torch.bfloat16, device=dev)
    a_flat[:rows_p] = a.view(rows_p, C)
    b = (torch.randn((N, 9 * C), device=dev, generator=g) * (9 * C) ** -0.5).to(torch.bfloat16)
    rows = n_img * H * W
    seg = [(i - 1) * wp + (j - 1) for i in range(3) for j in range(3)]
    outs, accs = [], []
    for halo in (False, True):
        out = torch.full((rows + 128, N), 3.0, dtype=torch.bfloat16, device=dev)
        acc = torch.zeros((ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64, device=dev)
        if halo:
            ops.conv3x3_halo(a_flat, n_img, H, W, C, b, out, gn_acc=acc)
        else:
            ops.gemm(a_flat, b, out, m_rows=rows_p, seg_off=seg, seg_k=C, remap=(hp, wp, 1, 1, H, W), gn_acc=acc,
                     gn_rows_per_img=H * W)
        torch.cuda.synchronize()
        outs.append(out.float().cpu().numpy())
        accs.append(acc.sum(0).cpu().numpy())
    ref = torch.nn.functional.conv2d(a[:, 1:-1, 1:-1].float().permute(0, 3, 1, 2),
                                     b.float().view(N, 3, 3, C).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1).reshape(rows, N)
    assert (outs[1][rows:] == 3.0).all(), "rows beyond the output stay untouched"
    err = np.abs(outs[1][:rows] - ref.cpu().numpy()).max()
    assert err <= 2.0 ** -7 * max(1.0, float(ref.abs().max())), err
    if C == 64:
        assert np.array_equal(outs[0][:rows], outs[1][:rows])
    else:
        d = np.abs(outs[0][:rows] - outs[1][:rows])
        assert d.max() <= 2.0 ** -7 * max(1.0, np.abs(outs[0]).max()) and (d > 0).mean() < 0.05
    np.testing.assert_allclose(accs[1], accs[0], rtol=1e-3 if C > 64 else 1e-12, atol=1e-3 if C > 64 else 1e-9)
