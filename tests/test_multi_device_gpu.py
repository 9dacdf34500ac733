"""One process, several devices -- the reference's host model (`jax.pmap` over the 8 devices of a box,
snap/trainer.py:452-464): every kernel attribute the library sets (dynamic shared memory opt-in, SM count) is per device.
Round 1 kept them in process-wide statics, so a second device launched its tcgen05 kernels with the default 48 KB limit.
The test touches device 1 FIRST, then device 0, and compares the results bit for bit.  Needs 2 GPUs (gpurun --gpus 2)."""
import numpy as np
import pytest
import torch

from util import F, bf16_np

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _run_all(dev):
    from snap_b200 import bev_mapper, configs, ops, params, pose_exhaustive_voting as pv, streetview_encoder as sve, synthetic, types
    from snap_b200.image_encoder import _WeightBank
    torch.cuda.set_device(dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(dev)
    out = {}
    # 1. the tcgen05 GEMM engine (200 KB of dynamic shared memory)
    rng = np.random.default_rng(1)
    a = t(bf16_np(rng.standard_normal((512, 256)))).to(torch.bfloat16)
    w = t(bf16_np(rng.standard_normal((256, 256)) * 0.1)).to(torch.bfloat16)
    y = torch.zeros((512, 256), dtype=torch.bfloat16, device=dev)
    ops.gemm(a, w, y, m_rows=512)
    out["gemm"] = y.float().cpu()
    # 2. the fused lift (226 KB)
    G, V, hw_img = 32, 2, (96, 128)
    hf, wf = 24, 32
    data = synthetic.make_tile(3, V, hw_img, G)
    mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
    xs, ys, zs = mapper.build_xyz_grid(data)
    Z = zs.shape[1]
    fp = params.round_to_bf16(params.init_mlp(rng, 257, (256, 128)))
    bank = _WeightBank(torch.device(dev))
    w0 = bank.add(fp["Dense_0"]["kernel"], False, 32)
    w1 = bank.add(fp["Dense_1"]["kernel"], False)
    bank.finalize(); bank.run()
    lp = sve.fill_lift_params(configs.streetview_encoder(), V, hf, wf, G, G, Z, 288)
    views = torch.from_numpy(sve.pack_views(data["camera"], data["T_view2scene"], 0, (4.0, 4.0))).to(dev)
    fimg = t(bf16_np(rng.standard_normal((1, V * hf * wf, 160)))).to(torch.bfloat16)
    plane = torch.zeros((1, G * G, 128), dtype=torch.bfloat16, device=dev)
    pvalid = torch.zeros((1, G * G), dtype=torch.uint8, device=dev)
    counter = torch.zeros(16, dtype=torch.int32, device=dev)
    scratch = torch.zeros(ops.lift_fused_batched_scratch_bytes(), dtype=torch.uint8, device=dev)
    ops.lift_fused_batched(lp, 1, views.view(1, -1), fimg, t(xs), t(ys), t(zs[:1]), bank.b_mats[w0], t(fp["Dense_0"]["kernel"][256]),
                           t(fp["Dense_0"]["bias"]), bank.b_mats[w1], t(fp["Dense_1"]["bias"]), plane, pvalid, counter, scratch)
    out["lift"] = plane.float().cpu()
    out["lift_valid"] = pvalid.cpu()
    # 3. the exhaustive correlation (xcorr_rows_kernel)
    Gx, R = 32, 8
    q = torch.nn.functional.normalize(t(rng.standard_normal((1, Gx, Gx, 32))), dim=-1).to(torch.bfloat16)
    m = torch.nn.functional.normalize(t(rng.standard_normal((1, Gx, Gx, 32))), dim=-1).to(torch.bfloat16)
    ones = torch.ones((1, Gx, Gx), dtype=torch.uint8, device=dev)
    out["xcorr"] = pv.exhaustive_pose_voting(types.FeaturePlane(q, ones), types.FeaturePlane(m, ones), R,
                                             types.Grid2D((Gx, Gx), 0.2)).cpu()
    # 4. round-2 kernels of the default encoder path: GroupNorm fused into a 1x1 conv (A_TGN1, 113 KB x 2 CTAs) and the halo
    #    3x3 conv (200 KB), each with its own per-device shared-memory opt-in
    n_img, H, W, Cc = 2, 24, 40, 256
    x = t(bf16_np(rng.standard_normal((n_img * H * W, Cc)))).to(torch.bfloat16)
    acc = torch.zeros((ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64, device=dev)
    ops.gn_stats(x, n_img, H * W, Cc, False, acc)
    wb = t(bf16_np(rng.standard_normal((64, Cc)) * 0.1)).to(torch.bfloat16)
    yc = torch.zeros((n_img * H * W + 128, 64), dtype=torch.bfloat16, device=dev)
    ops.conv_gn(x, n_img, H, W, Cc, acc, torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev), wb, yc)
    out["conv_gn"] = yc.float().cpu()
    a3 = torch.zeros((n_img, H + 2, W + 2, 64), dtype=torch.bfloat16, device=dev)
    a3[:, 1:-1, 1:-1] = t(bf16_np(rng.standard_normal((n_img, H, W, 64)))).to(torch.bfloat16)
    a3f = torch.zeros((n_img * (H + 2) * (W + 2) + 256, 64), dtype=torch.bfloat16, device=dev)
    a3f[: n_img * (H + 2) * (W + 2)] = a3.view(-1, 64)
    w3 = t(bf16_np(rng.standard_normal((64, 9 * 64)) * 0.05)).to(torch.bfloat16)
    y3 = torch.zeros((n_img * H * W + 128, 64), dtype=torch.bfloat16, device=dev)
    ops.conv3x3_halo(a3f, n_img, H, W, 64, w3, y3)
    out["conv3x3_halo"] = y3.float().cpu()
    torch.cuda.synchronize(dev)
    return out


def test_kernels_run_on_a_second_device_of_the_same_process():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process (gpurun --gpus 2)")
    second = _run_all("cuda:1")       # device 1 first: nothing was configured by an earlier launch on device 0
    first = _run_all("cuda:0")
    torch.cuda.set_device(0)
    assert float(second["gemm"].abs().sum()) > 0 and int(second["lift_valid"].sum()) > 0
    for k in first:
        a, b = first[k], second[k]
        assert torch.equal(torch.nan_to_num(a.float(), neginf=-1e30), torch.nan_to_num(b.float(), neginf=-1e30)), k
