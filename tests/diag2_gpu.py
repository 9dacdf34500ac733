import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import F, bf16_np, rd_bf16, rel_l2
from snap_b200 import ops, params
from snap_b200.image_encoder import _WeightBank
from oracle import layers as olayers

_t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F))
dev = "cuda"
rng = np.random.default_rng(11)
rd = lambda a: rd_bf16(_t(a)).numpy()

for N in (900, 34560, 245760):
    print("== N", N)
    stats = np.zeros((N, 288), F)
    valid = rng.random(N) < 0.03
    stats[valid, :257] = bf16_np(rng.standard_normal((int(valid.sum()), 257)))
    fp = params.round_to_bf16(params.perturb_affine(rng, params.init_mlp(rng, 257, (256, 128))))
    bank = _WeightBank(torch.device(dev))
    w0 = bank.add(fp["Dense_0"]["kernel"], False, 32)
    w1 = bank.add(fp["Dense_1"]["kernel"], False)
    bank.finalize(); bank.run()
    sd = _t(stats).to(torch.bfloat16).to(dev)
    hid32 = torch.zeros((N, 256), device=dev)
    ops.gemm(sd, bank.b_mats[w0], hid32, m_rows=N, seg_k=288)
    hid = torch.zeros((N, 256), dtype=torch.bfloat16, device=dev)
    ops.gemm(sd, bank.b_mats[w0], hid, m_rows=N, seg_k=288, bias=_t(fp["Dense_0"]["bias"]).to(dev), relu=True)
    vol = torch.zeros((N, 128), dtype=torch.bfloat16, device=dev)
    vd = torch.from_numpy(valid.astype(np.uint8)).to(dev)
    ops.gemm(hid, bank.b_mats[w1], vol, m_rows=N, bias=_t(fp["Dense_1"]["bias"]).to(dev), row_mask=vd)
    torch.cuda.synchronize()
    W0 = np.zeros((288, 256), F); W0[:257] = fp["Dense_0"]["kernel"]
    ref32 = stats @ W0
    g = hid32.cpu().numpy()
    err_rows = np.abs(g - ref32).max(1)
    bad = err_rows > 1e-3 * np.abs(ref32).max()
    print(f" GEMM1 f32: bad rows {bad.sum()} / {N}; rel_l2 {rel_l2(g, ref32):.3e}; first bad rows {np.nonzero(bad)[0][:10]}, tiles {np.unique(np.nonzero(bad)[0] // 128)[:20]}")
    wdev = bank.b_mats[w0].float().cpu().numpy()
    print("  W0 device == host:", np.array_equal(wdev[:, :257], fp["Dense_0"]["kernel"].T), " pad zero:", not wdev[:, 257:].any())
    h_ref = np.maximum(rd(rd(ref32) + fp["Dense_0"]["bias"]), 0)
    gh = hid.float().cpu().numpy()
    print(f" hid bf16: rel_l2 {rel_l2(gh, h_ref):.3e}, mismatches {(gh != h_ref).mean():.4%}")
    v_ref = rd(rd(gh @ fp["Dense_1"]["kernel"]) + fp["Dense_1"]["bias"]) * valid[:, None]
    gv = vol.float().cpu().numpy()
    er = np.abs(gv - v_ref).max(1)
    badv = er > 2e-2 * np.abs(v_ref).max()
    print(f" vol bf16 (from GPU hid): rel_l2 {rel_l2(gv, v_ref):.3e}, mismatches {(gv != v_ref).mean():.4%}, bad rows {badv.sum()} first {np.nonzero(badv)[0][:10]}")
    print(f"   invalid rows all zero: {not gv[~valid].any()}")
