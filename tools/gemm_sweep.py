"""Per-layer micro-benchmark of the encoder's GEMM / GroupNorm-apply launches at the bench batch (16 images of
512x672): time, algorithmic bytes, GB/s and TFLOP/s per launch, with epilogue features toggled, so that the
kernels can be placed on their rooflines one by one.  Run on a B200: `python tools/gemm_sweep.py`."""
import json
import sys

import torch

sys.path.insert(0, ".")
from snap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
NIMG = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bf(rows, cols):
    return (torch.randn((rows, cols), device=dev, dtype=torch.float32) * 0.5).to(torch.bfloat16)


def case(name, h, w, k, n, taps=1, res=False, gn=True, gn_relu=False, bn=0):
    rows = NIMG * h * w
    acc = torch.zeros((ops.GN_REPLICAS, NIMG, 32, 2), dtype=torch.float64, device=dev)
    acc2 = torch.zeros_like(acc)
    out = torch.zeros((rows + 128, n), dtype=torch.bfloat16, device=dev)
    r = bf(rows + 128, n) if res else None
    if taps == 1:
        a = bf(rows + 128, k)
        b = bf(max(n, 16), k)
        fn = lambda: ops.gemm(a, b, out, m_rows=rows, residual=r, bn=bn, gn_acc=acc if gn else None,
                              gn_acc_relu=acc2 if (gn and gn_relu) else None, gn_rows_per_img=h * w)
        a_bytes = rows * k * 2
    else:
        hp, wp = h + 2, w + 2
        a = bf(NIMG * hp * wp + 4 * wp + 256, k)
        b = bf(max(n, 16), 9 * k)
        seg = [(i - 1) * wp + (j - 1) for i in range(3) for j in range(3)]
        seg = [s + wp + 1 for s in seg]  # keep offsets non-negative for the standalone buffer
        fn = lambda: ops.gemm(a, b, out, m_rows=NIMG * hp * wp, seg_off=seg, seg_k=k, remap=(hp, wp, 0, 0, h, w),
                              bn=bn, gn_acc=acc if gn else None, gn_rows_per_img=h * w)
        a_bytes = NIMG * hp * wp * k * 2
    ms = timeit(fn)
    byt = a_bytes + rows * n * 2 * (2 if res else 1) + n * k * taps * 2
    fl = 2.0 * rows * n * k * taps
    print(f"{name:34s} M={rows:7d} K={k * taps:5d} N={n:5d} res={int(res)} gn={int(gn)}{'+r' if gn_relu else '  '} bn={bn:3d}"
          f"  {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  {fl / ms / 1e9:7.0f} TF/s", flush=True)
    return dict(name=name, ms=ms, gbps=byt / ms / 1e6, tflops=fl / ms / 1e9)


def fused_case(name, h, w, k, n, res=False, gn=True, gn_relu=False):
    """conv_gn (A_TGN1): GroupNorm + ReLU of the raw input inside the GEMM; compare with gn_apply + gemm."""
    rows = NIMG * h * w
    x = bf(rows + 128, k)
    b = bf(max(n, 16), k)
    acc_in = torch.zeros((ops.GN_REPLICAS, NIMG, 32, 2), dtype=torch.float64, device=dev)
    ops.gn_stats(x, NIMG, h * w, k, False, acc_in)
    acc = torch.zeros_like(acc_in)
    acc2 = torch.zeros_like(acc_in)
    out = torch.zeros((rows + 128, n), dtype=torch.bfloat16, device=dev)
    r = bf(rows + 128, n) if res else None
    sc = torch.ones(k, device=dev)
    bi = torch.zeros(k, device=dev)
    ms = timeit(lambda: ops.conv_gn(x, NIMG, h, w, k, acc_in, sc, bi, b, out, residual=r, gn_acc=acc if gn else None,
                                    gn_acc_relu=acc2 if (gn and gn_relu) else None))
    byt = rows * k * 2 + rows * n * 2 * (2 if res else 1) + n * k * 2
    fl = 2.0 * rows * n * k
    print(f"{name:34s} M={rows:7d} K={k:5d} N={n:5d} res={int(res)} gn={int(gn)}{'+r' if gn_relu else '  '} FUSED "
          f"  {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  {fl / ms / 1e9:7.0f} TF/s", flush=True)
    return dict(name=name + " fused", ms=ms, gbps=byt / ms / 1e6, tflops=fl / ms / 1e9)


def halo_case(name, h, w, k, n, gn=True):
    rows = NIMG * h * w
    hp, wp = h + 2, w + 2
    if not ops.conv3x3_halo_supported(k, n, w):
        print(f"{name:34s} unsupported by the halo kernel", flush=True)
        return dict(name=name + " halo", ms=None)
    a = bf(NIMG * hp * wp + 256, k)
    b = bf(n, 9 * k)
    acc = torch.zeros((ops.GN_REPLICAS, NIMG, 32, 2), dtype=torch.float64, device=dev)
    out = torch.zeros((rows + 128, n), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.conv3x3_halo(a, NIMG, h, w, k, b, out, gn_acc=acc if gn else None))
    byt = NIMG * hp * wp * k * 2 + rows * n * 2 + 9 * n * k * 2
    fl = 2.0 * rows * n * k * 9
    print(f"{name:34s} M={rows:7d} K={k * 9:5d} N={n:5d} gn={int(gn)} HALO "
          f"  {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s  {fl / ms / 1e9:7.0f} TF/s", flush=True)
    return dict(name=name + " halo", ms=ms, gbps=byt / ms / 1e6, tflops=fl / ms / 1e9)


def gn_case(name, h, w, c, layout):
    rows = NIMG * h * w
    x = bf(rows + 128, c)
    acc = torch.zeros((ops.GN_REPLICAS, NIMG, 32, 2), dtype=torch.float64, device=dev)
    ops.gn_stats(x, NIMG, h * w, c, False, acc)
    sc = torch.ones(c, device=dev)
    bi = torch.zeros(c, device=dev)
    orow = rows if layout == ops.LAYOUT_DENSE else NIMG * (h + 2) * (w + 2)
    out = torch.zeros((orow + 256, c), dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.gn_apply(x, NIMG, h, w, c, acc, sc, bi, False, True, layout, out))
    byt = rows * c * 2 + orow * c * 2
    print(f"{name:34s} rows={rows:7d} C={c:5d}  {ms * 1e3:8.1f} us  {byt / ms / 1e6:7.0f} GB/s", flush=True)
    return dict(name=name, ms=ms, gbps=byt / ms / 1e6)


res = []
S = [(128, 168), (64, 84), (32, 42), (16, 21)]
print("== 1x1 convs ==")
res.append(case("s1.conv1 (256->64)", *S[0], 256, 64))
res.append(case("s1.conv1 (256->64) no-gn", *S[0], 256, 64, gn=False))
res.append(case("s1.conv3 (64->256)+res", *S[0], 64, 256, res=True))
res.append(case("s1.conv3 (64->256)+res no-gn", *S[0], 64, 256, res=True, gn=False))
res.append(case("s1.conv3 (64->256) no-res no-gn", *S[0], 64, 256, gn=False))
res.append(case("s1.conv3 (64->256)+res bn=256", *S[0], 64, 256, res=True, bn=256))
res.append(case("s1.conv3 (64->256)+res bn=64", *S[0], 64, 256, res=True, bn=64))
res.append(case("s1.conv3 +res gn+relu", *S[0], 64, 256, res=True, gn_relu=True))
res.append(case("s2.conv1 unit1 (256->128)", *S[0], 256, 128))
res.append(case("s2.conv1 (512->128)", *S[1], 512, 128))
res.append(case("s2.conv3 (128->512)+res", *S[1], 128, 512, res=True))
res.append(case("s2.conv3 (128->512)+res no-gn", *S[1], 128, 512, res=True, gn=False))
res.append(case("s2.proj (256->512)", *S[1], 256, 512, gn=False))
res.append(case("s3.conv1 (1024->256)", *S[2], 1024, 256))
res.append(case("s3.conv3 (256->1024)+res", *S[2], 256, 1024, res=True))
res.append(case("s4.conv1 (2048->512)", *S[3], 2048, 512))
res.append(case("s4.conv3 (512->2048)+res", *S[3], 512, 2048, res=True))
print("== fused GroupNorm -> 1x1 convs (A_TGN1) ==")
res.append(fused_case("s1.conv1 (256->64)", *S[0], 256, 64))
res.append(fused_case("s1.conv3 (64->256)+res", *S[0], 64, 256, res=True))
res.append(fused_case("s1.conv3 +res gn+relu", *S[0], 64, 256, res=True, gn_relu=True))
res.append(fused_case("s1.proj (64->256)", *S[0], 64, 256, gn=False))
res.append(fused_case("s2.conv1 unit1 (256->128)", *S[0], 256, 128))
res.append(fused_case("s2.conv1 (512->128)", *S[1], 512, 128))
res.append(fused_case("s2.conv3 (128->512)+res", *S[1], 128, 512, res=True))
res.append(fused_case("s3.conv1 (1024->256)", *S[2], 1024, 256))
res.append(fused_case("s3.conv3 (256->1024)+res", *S[2], 256, 1024, res=True))
res.append(fused_case("s4.conv1 (2048->512)", *S[3], 2048, 512))
res.append(fused_case("s4.conv3 (512->2048)+res", *S[3], 512, 2048, res=True))
res.append(fused_case("fpn s1 (256->128)+res", *S[0], 256, 128, res=True, gn=False))
res.append(fused_case("fpn s2 (512->128)+res", *S[1], 512, 128, res=True, gn=False))
print("== 3x3 convs ==")
res.append(case("s1.conv2 3x3 64", *S[0], 64, 64, taps=9))
res.append(case("s1.conv2 3x3 64 no-gn", *S[0], 64, 64, taps=9, gn=False))
res.append(case("s2.conv2 3x3 128", *S[1], 128, 128, taps=9))
res.append(case("s3.conv2 3x3 256", *S[2], 256, 256, taps=9))
res.append(case("s3.conv2 3x3 256 bn=256", *S[2], 256, 256, taps=9, bn=256))
res.append(case("s4.conv2 3x3 512", *S[3], 512, 512, taps=9))
print("== 3x3 convs, halo kernel ==")
res.append(halo_case("s1.conv2 3x3 64", *S[0], 64, 64))
res.append(halo_case("s1.conv2 3x3 64 no-gn", *S[0], 64, 64, gn=False))
res.append(halo_case("s2.conv2 3x3 128", *S[1], 128, 128))
res.append(halo_case("s3.conv2 3x3 256", *S[2], 256, 256))
print("== GroupNorm apply ==")
res.append(gn_case("gn s1 256 dense", *S[0], 256, ops.LAYOUT_DENSE))
res.append(gn_case("gn s1 64 dense", *S[0], 64, ops.LAYOUT_DENSE))
res.append(gn_case("gn s1 64 padded", *S[0], 64, ops.LAYOUT_PADDED))
res.append(gn_case("gn s2 512 dense", *S[1], 512, ops.LAYOUT_DENSE))
res.append(gn_case("gn s2 128 padded", *S[1], 128, ops.LAYOUT_PADDED))
res.append(gn_case("gn s3 1024 dense", *S[2], 1024, ops.LAYOUT_DENSE))
res.append(gn_case("gn s3 256 padded", *S[2], 256, ops.LAYOUT_PADDED))
res.append(gn_case("gn s4 2048 dense", *S[3], 2048, ops.LAYOUT_DENSE))
json.dump(res, open("gpurun_out/gemm_sweep.json", "w"), indent=1)
