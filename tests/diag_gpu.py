"""One-shot GPU diagnostics (not a pytest): locates where CUDA and oracle diverge."""
import sys, os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from util import F, bf16_np, rd_bf16, rel_l2, to_oracle_geometry
from snap_b200 import ops, configs, params, image_encoder, streetview_encoder as sve, bev_mapper, synthetic, types
from oracle import resnet as ores, image_encoder as oie, bev_mapper as obm, grids as ogrids, streetview_encoder as osv, layers as olayers

_t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F))
dev = "cuda"


def mismatch(a, b, name):
    a = np.asarray(a, F); b = np.asarray(b, F)
    ne = a != b
    print(f"  {name}: rel_l2 {rel_l2(a, b):.3e}, mismatching elements {ne.mean():.4%}, max abs {np.abs(a-b).max():.4g} (scale {np.abs(b).max():.4g})")


print("== 1. GEMM exactness vs float64")
for (M, K, N) in [(512, 64, 64), (512, 576, 64), (512, 2048, 256), (512, 4608, 512)]:
    g = torch.Generator().manual_seed(K)
    a = torch.randn((M, K), generator=g).to(torch.bfloat16); b = torch.randn((N, K), generator=g).to(torch.bfloat16)
    out = torch.zeros((M, N), device=dev)
    ops.gemm(a.to(dev), b.to(dev), out); torch.cuda.synchronize()
    ref64 = a.double() @ b.double().T
    ref32 = (a.float() @ b.float().T)
    e = (out.cpu().double() - ref64).abs().max().item() / ref64.abs().max().item()
    e32 = (ref32.double() - ref64).abs().max().item() / ref64.abs().max().item()
    print(f"  K={K}: GPU max err/scale {e:.3e}; torch-CPU fp32 {e32:.3e}")

print("== 2. weight standardisation: bf16 mismatches vs oracle")
rng = np.random.default_rng(0)
bank = image_encoder._WeightBank(torch.device(dev))
shapes = [(7, 7, 3, 64), (1, 1, 64, 256), (3, 3, 128, 128), (1, 1, 1024, 2048)]
ws = [params.round_to_bf16({"k": params.lecun_normal(rng, s)})["k"] for s in shapes]
ids = [bank.add(w, True, 32) for w in ws]
bank.finalize(); bank.run(); torch.cuda.synchronize()
for i, (s, w) in enumerate(zip(shapes, ws)):
    ref = rd_bf16(ores.std_kernel(_t(w))).reshape(-1, s[-1]).numpy().T
    K = ref.shape[1]
    mismatch(bank.b_mats[i].float().cpu().numpy()[: s[-1], :K], ref, f"std {s}")

print("== 3. GroupNorm: bf16 mismatches vs oracle (bf16 mode)")
for (n, h, w, c) in [(2, 16, 24, 64), (2, 16, 24, 256), (2, 4, 6, 1024)]:
    x = bf16_np(rng.standard_normal((n, h, w, c)) * 3 + 1.0)
    scale, bias = bf16_np(1 + 0.3 * rng.standard_normal(c)), bf16_np(0.2 * rng.standard_normal(c))
    ref = torch.relu(ores.group_norm(_t(x), _t(scale), _t(bias), rd_bf16)).numpy()
    xd = _t(x).to(torch.bfloat16).to(dev)
    stats = torch.zeros((n, 32, 2), device=dev)
    wsb = torch.zeros(ops.gn_workspace_bytes(n, h * w) // 4 + 16, device=dev)
    out = torch.zeros((n, h, w, c), dtype=torch.bfloat16, device=dev)
    ops.gn_stats(xd, n, h * w, c, False, stats, wsb)
    ops.gn_apply(xd, n, h, w, c, stats, _t(scale).to(dev), _t(bias).to(dev), False, True, ops.LAYOUT_DENSE, out)
    torch.cuda.synchronize()
    xg = _t(x).reshape(n, h, w, 32, c // 32)
    mean = xg.mean(dim=[1, 2, 4]); var = ((xg - xg.mean(dim=[1, 2, 4], keepdim=True)) ** 2).mean(dim=[1, 2, 4])
    st = stats.cpu()
    print(f"  C={c}: mean max rel err {((st[..., 0] - mean).abs() / (mean.abs() + 1e-6)).max():.2e}, rstd max rel err {((st[..., 1] - 1 / torch.sqrt(var + 1e-5)).abs() * torch.sqrt(var + 1e-5)).max():.2e}")
    mismatch(out.float().cpu().numpy(), ref, f"gn C={c}")

print("== 4. one residual unit, intermediate by intermediate (teacher forced)")
cfg = configs.image_encoder()
rng = np.random.default_rng(2)
p = params.round_to_bf16(params.perturb_affine(rng, params.init_image_encoder(rng, cfg)))
img = rng.random((2, 40, 72, 3), dtype=F)
tt = lambda tree: {k: (tt(v) if isinstance(v, dict) else _t(v)) for k, v in tree.items()}
trace = []
oie.image_encoder(_t(img), tt(p), False, rd_bf16, trace)
enc = image_encoder.ImageEncoder(cfg)
plan = enc.plan(p, 2, 40, 72, torch.device(dev))
plan.bank.run()
for ui in (1, 3):
    u = plan.units[ui]
    xin, yout = trace[ui + 1]
    pu = tt(p["encoder"])[f"block{1 if ui < 3 else 2}"][f"unit{(ui % 3) + 1:02d}"]
    rows_in = xin.numel() // xin.shape[-1]
    xd = torch.zeros((max(128, -(-rows_in // 128) * 128), xin.shape[-1]), dtype=torch.bfloat16, device=dev)
    xd[:rows_in] = xin.reshape(rows_in, -1).to(torch.bfloat16).to(dev)
    out = plan.run_unit(u, xd); torch.cuda.synchronize()
    rd = rd_bf16
    a1 = torch.relu(ores.group_norm(xin, pu["gn1"]["scale"], pu["gn1"]["bias"], rd))
    y1 = ores.conv(a1, ores.std_kernel(pu["conv1"]["kernel"], rd), rd=rd)
    a2 = torch.relu(ores.group_norm(y1, pu["gn2"]["scale"], pu["gn2"]["bias"], rd))
    y2 = ores.conv(a2, ores.std_kernel(pu["conv2"]["kernel"], rd), stride=u["stride"], padding=1, rd=rd)
    a3 = torch.relu(ores.group_norm(y2, pu["gn3"]["scale"], pu["gn3"]["bias"], rd))
    y3 = ores.conv(a3, ores.std_kernel(pu["conv3"]["kernel"], rd), rd=rd)
    rows_out = y2.numel() // y2.shape[-1]
    print(f" unit {ui}: stride {u['stride']}")
    # buffers after the run: buf_a holds a3, buf_y holds y2, u['a2'] holds a2, u['out'] the output
    mismatch(plan._view(plan.buf_y, rows_out, u["nmid"])[:rows_out].float().cpu().numpy(), y2.reshape(rows_out, -1).numpy(), "y2 (conv2 out)")
    mismatch(plan._view(plan.buf_a, rows_out, u["nmid"])[:rows_out].float().cpu().numpy(), a3.reshape(rows_out, -1).numpy(), "a3 (gn3 out)")
    if u["stride"] == 1:
        a2g = u["a2"][: 2 * (u["h"] + 2) * (u["w"] + 2)].view(2, u["h"] + 2, u["w"] + 2, -1)[:, 1:-1, 1:-1].float().cpu().numpy()
        mismatch(a2g, a2.numpy(), "a2 (gn2 out)")
    mismatch(out[:rows_out].float().cpu().numpy(), yout.reshape(rows_out, -1).numpy(), "unit out")
    # GPU conv1 on the ORACLE a1 -> isolates conv1
    a1d = torch.zeros_like(xd[:, : a1.shape[-1]].contiguous()) if a1.shape[-1] <= xd.shape[1] else None
    a1d = torch.zeros((xd.shape[0], a1.shape[-1]), dtype=torch.bfloat16, device=dev)
    a1d[:rows_in] = a1.reshape(rows_in, -1).to(torch.bfloat16).to(dev)
    y1g = torch.zeros((xd.shape[0], u["nmid"]), dtype=torch.bfloat16, device=dev)
    ops.gemm(a1d, plan.bank.b_mats[u["w1"]], y1g, m_rows=rows_in); torch.cuda.synchronize()
    mismatch(y1g[:rows_in].float().cpu().numpy(), y1.reshape(rows_in, -1).numpy(), "conv1 on oracle a1")
    wref = rd(ores.std_kernel(pu["conv1"]["kernel"])).reshape(-1, u["nmid"]).numpy().T
    mismatch(plan.bank.b_mats[u["w1"]].float().cpu().numpy()[:, : wref.shape[1]], wref, "conv1 std weights")

print("== 5. lift statistics per voxel")
for fisheye, V in [(False, 3), (False, 1)]:
    G, hw_img = 24, (64, 96)
    data = synthetic.make_tile(5, V, hw_img, G, fisheye=fisheye)
    grid = types.Grid2D((G, G), 0.2)
    mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), grid)
    xs, ys, zs = mapper.build_xyz_grid(data)
    hf, wf = 16, 24
    rng = np.random.default_rng(11)
    scfg = configs.streetview_encoder()
    Z = zs.shape[1]; N = G * G * Z
    fimg_np = bf16_np(rng.standard_normal((V, hf, wf, 160)))
    lp = sve.fill_lift_params(scfg, V, hf, wf, G, G, Z, 288)
    views = torch.from_numpy(sve.pack_views(data["camera"], data["T_view2scene"], 0, (4.0, 4.0))).to("cuda")
    stats = torch.zeros((N, 288), dtype=torch.bfloat16, device=dev)
    valid = torch.zeros(N, dtype=torch.uint8, device=dev)
    ops.lift_gather_pool(lp, views, _t(fimg_np).to(torch.bfloat16).to(dev), _t(xs).to(dev), _t(ys).to(dev), _t(zs[0]).to(dev), stats, valid)
    torch.cuda.synchronize()
    ocam, oT = to_oracle_geometry(data, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz, _ = obm.build_xyz_query(ogrids.Grid2D((G, G), 0.2), oT.t)
    rdn = obm.np_rd(rd_bf16)
    p2d, vis, depth, _ = osv.project_points_to_views(oT, ocam, xyz.reshape(-1, 3))
    f_proj = rdn(osv.interpolate_views_all(fimg_np, p2d))
    scores = rdn(osv.interpolate_depth_score(f_proj[..., 128:], depth))
    ostats, ovalid = osv.pool_multiview_features(f_proj[..., :128], vis, scores, False, True, rd=rdn)
    ostats = rdn(ostats)
    g = stats.float().cpu().numpy()[:, :257]
    v = ovalid
    print(f" fisheye={fisheye} V={V}: valid {v.sum()}, stats cols>=257 nonzero: {bool(stats.float().cpu().numpy()[:, 257:].any())}")
    mismatch(g[v][:, :128], ostats[v][:, :128], "mean")
    mismatch(g[v][:, 128:256], ostats[v][:, 128:256], "var")
    mismatch(g[v][:, 256], ostats[v][:, 256], "score_max")
    err = np.abs(g[v] - ostats[v]).max(1)
    worst = np.argsort(-err)[:5]
    idx = np.nonzero(v)[0][worst]
    for k, i in zip(worst, idx):
        print(f"   voxel {i}: err {err[k]:.4g} vis {vis[i].astype(int)} p2d {p2d[i].round(2).tolist()} depth {depth[i].round(2).tolist()} scores {scores[i].tolist()} gpu smax {g[i, 256]:.4g}")
