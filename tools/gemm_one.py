"""One encoder GEMM configuration, a few launches (ncu target): python tools/gemm_one.py conv3|conv1|conv2"""
import sys

import torch

sys.path.insert(0, ".")
from snap_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "conv3"
NIMG, h, w = 16, 128, 168
rows = NIMG * h * w
bf = lambda r, c: (torch.randn((r, c), device=dev) * 0.5).to(torch.bfloat16)
acc = torch.zeros((ops.GN_REPLICAS, NIMG, 32, 2), dtype=torch.float64, device=dev)
if which == "conv3":
    a, b, out, r = bf(rows + 128, 64), bf(256, 64), bf(rows + 128, 256), bf(rows + 128, 256)
    fn = lambda: ops.gemm(a, b, out, m_rows=rows, residual=r, gn_acc=acc, gn_rows_per_img=h * w)
elif which == "conv1":
    a, b, out = bf(rows + 128, 256), bf(64, 256), bf(rows + 128, 64)
    fn = lambda: ops.gemm(a, b, out, m_rows=rows, gn_acc=acc, gn_rows_per_img=h * w)
elif which == "root":  # implicit 7x7/2 root conv + max pool + statistics of 16 images of 480x640
    import numpy as np
    from snap_b200 import configs, image_encoder, params
    cfg = configs.image_encoder()
    p_ = params.init_image_encoder(np.random.default_rng(0), cfg)
    plan = image_encoder.ImageEncoder(cfg).plan(p_, NIMG, 480, 640, dev)
    imgs = torch.rand((NIMG, 480, 640, 3), device=dev)
    plan.bank.run()
    fn = lambda: plan.run_root(imgs)
elif which == "proj160":  # proj MLP: Dense 128 -> 160 with bias over 16 x 120 x 160 texels
    rows_p = NIMG * 120 * 160
    a, b, out = bf(rows_p + 128, 128), bf(160, 128), bf(rows_p + 128, 160)
    bias = torch.randn(160, device=dev)
    fn = lambda: ops.gemm(a, b, out, m_rows=rows_p, bias=bias)
elif which in ("fconv1", "fconv3"):  # fused GroupNorm -> 1x1 conv (A_TGN1)
    k, n = (256, 64) if which == "fconv1" else (64, 256)
    x, b, out = bf(rows + 128, k), bf(n, k), bf(rows + 128, n)
    r = bf(rows + 128, n) if which == "fconv3" else None
    acc_in = torch.zeros_like(acc)
    ops.gn_stats(x, NIMG, h * w, k, False, acc_in)
    sc, bi = torch.ones(k, device=dev), torch.zeros(k, device=dev)
    fn = lambda: ops.conv_gn(x, NIMG, h, w, k, acc_in, sc, bi, b, out, residual=r, gn_acc=acc)
else:
    hp, wp = h + 2, w + 2
    a, b, out = bf(NIMG * hp * wp + 4 * wp + 256, 64), bf(64, 576), bf(rows + 128, 64)
    seg = [i * wp + j for i in range(3) for j in range(3)]
    fn = lambda: ops.gemm(a, b, out, m_rows=NIMG * hp * wp, seg_off=seg, seg_k=64, remap=(hp, wp, 0, 0, h, w),
                          gn_acc=acc, gn_rows_per_img=h * w)
for _ in range(4):
    fn()
torch.cuda.synchronize()
