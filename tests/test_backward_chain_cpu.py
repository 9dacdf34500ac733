"""Capstone of the CPU-verified backward plans: the three product plans chained exactly as a config-4 training step with
frozen image encoders will chain them --

    NLL of the sampled poses -> `LocalizerLossBackward` -> cotangent of the map's `bev_matching`
                             -> `MatchingHeadBackward` -> cotangent of the street-view plane
                             -> `LiftBackward.scene_backward` -> gradients of fusion_mlp / proj_mlp, encoder-feature cotangent

on the emulated operator layer, against ONE torch autograd graph of the oracle chain
encoder features -> proj MLP -> lift -> fusion MLP -> vertical max -> matching head -> similarities -> pose scores -> NLL.
This pins the interfaces between the plans (layouts, dtypes, masks), which the per-plan tests cannot."""
import numpy as np
import torch

from lift_torch_ref import chain_forward
from loc_torch_ref import nll, pose_scores, pose_uv
from ops_emulation import emulated_ops, make_lift_emulation
from util import F, bf16_np, rd_bf16, to_oracle_geometry


def test_loss_to_lift_backward_chain_matches_autograd():
    from oracle import bev_mapper as obm, grids as ogrids, streetview_encoder as osv
    from snap_b200 import configs, localizer_train, params, pose_estimation, streetview_encoder as sve, streetview_train, synthetic
    G, V, hw, hf, wf, D = 16, 3, (64, 96), 16, 24, 32
    cell = 0.2
    rng = np.random.default_rng(17)
    data = synthetic.make_tile(6, V, hw, G, spacing=0.5, same_side=True)
    ocam, oT = to_oracle_geometry(data, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz, _ = obm.build_xyz_query(ogrids.Grid2D((G, G), cell), oT.t)
    Z = xyz.shape[2]
    cells, N_vox = G * G, G * G * Z
    p2d, vis, depth, _ = osv.project_points_to_views(oT, ocam, xyz.reshape(-1, 3))
    svp = params.round_to_bf16(params.perturb_affine(rng, {"proj_mlp": params.init_mlp(rng, 128, (160,)),
                                                           "fusion_mlp": params.init_mlp(rng, 257, (256, 128))}))
    mp = {"kernel": bf16_np(rng.standard_normal((128, D)) * 0.2), "bias": bf16_np(rng.standard_normal(D) * 0.05)}
    enc = bf16_np(rng.standard_normal((V * hf * wf, 128)))
    # query side (constants here: the query mapper's chain is the same code) and sampled poses
    Nq, P1, temp = 24, 48, 0.2
    fq = bf16_np(rng.standard_normal((Nq, D)) / np.sqrt(D) * 2)
    valid_q = rng.random(Nq) < 0.85
    q_xy = np.stack([rng.uniform(0.2, 1.5, Nq), rng.uniform(-0.8, 0.8, Nq)], -1).astype(F)
    poses = np.stack([rng.uniform(-0.5, 0.5, P1), rng.uniform(0.2, 1.5, P1), rng.uniform(0.8, 2.4, P1)], -1).astype(F)

    # ---- reference: one autograd graph ---------------------------------------------------------------------------------
    tp, x, t = chain_forward(svp, enc, p2d, vis, depth, V, hf, wf, cells, Z, rd_bf16)
    K, bias = torch.from_numpy(mp["kernel"]).requires_grad_(True), torch.from_numpy(mp["bias"]).requires_grad_(True)
    T = torch.tensor(temp, requires_grad=True)
    pvalid = t["plane_valid"]
    y = rd_bf16(rd_bf16(t["plane"] @ K) + bias)                                             # bev_mapper.py:284-287
    nrm = y.norm(dim=-1, keepdim=True)
    fm = torch.where(pvalid[:, None] & (nrm >= 1e-5), rd_bf16(y / nrm.clamp(min=1e-30)), torch.zeros(()))   # :288-291
    sim = torch.relu(rd_bf16(torch.from_numpy(fq) @ fm.T))                                  # bev_localizer.py:157-159
    w = 1.0 / max(int(valid_q.sum()), 1)
    sc = pose_scores((sim * torch.exp(T) * w).reshape(Nq, G, G), pose_uv(poses, q_xy, cell), valid_q, pvalid.numpy().reshape(G, G), True)
    nll(sc, np.zeros(P1, bool)).backward()
    assert pvalid.float().mean() > 0.05 and float(sc.detach().abs().max()) > 0

    # ---- the product's plans, chained ----------------------------------------------------------------------------------
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    nump = lambda a: a.detach().numpy()
    lp = sve.fill_lift_params(configs.streetview_encoder(), V, hf, wf, G, G, Z, 288)
    scale = float(np.exp(F(temp)))
    maps = pose_estimation.SimilarityMaps(sim=bf(nump(sim))[None], scale=scale,
                                          point_scale=torch.from_numpy(np.where(valid_q, scale * w, 0).astype(F))[None],
                                          row_cdf=None, row_max=None, chunk_sum=None, row_sum=None, H=G, W=G)
    pv_u8 = torch.from_numpy(nump(pvalid).astype(np.uint8))
    with emulated_ops(make_lift_emulation(p2d, vis, depth)):
        dev = torch.device("cpu")
        loc = localizer_train.LocalizerLossBackward(dev)
        _, dfm, dtemp = loc.backward(maps, bf(fq)[None], bf(nump(fm)).reshape(1, G, G, D), torch.from_numpy(q_xy), pv_u8[None],
                                     torch.from_numpy(poses)[None], sc.detach()[None], cell, True, True)
        mh = streetview_train.MatchingHeadBackward(mp, dev)
        dplane = mh.backward(bf(nump(t["plane"])), pv_u8, dfm[0].to(torch.bfloat16).contiguous())
        lb = streetview_train.LiftBackward(svp, dev)
        lb.zero_grads()
        dcrop = lb.scene_backward(lp, None, bf(nump(t["fimg"])), bf(nump(t["crop"])), None, None, None, bf(nump(t["vol"])),
                                  torch.from_numpy(vis.any(-1).astype(np.uint8)), dplane[:cells].contiguous())
        got = lb.grads_tree()
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    errs = {"matching_proj/kernel": rel(mh.g["kernel"].numpy(), K.grad.numpy()),
            "matching_proj/bias": rel(mh.g["bias"].numpy(), bias.grad.numpy()),
            "temperature": abs(float(dtemp.sum()) - float(T.grad)) / (abs(float(T.grad)) + 1e-30),
            "encoder features": rel(dcrop[: V * hf * wf].float().numpy(), x.grad.numpy())}
    for k in ("proj_mlp", "fusion_mlp"):
        for n, d in tp[k].items():
            for a, leaf in d.items():
                assert float(leaf.grad.norm()) > 1e-6, (k, n, a)
                errs[f"{k}/{n}/{a}"] = rel(got[k][n][a], leaf.grad.numpy())
    print({k: round(float(v), 4) for k, v in errs.items()})
    # bf16 cotangents between the plans (d f_m, dplane) against fp32 cotangents in autograd: a few 1e-3
    assert max(errs.values()) < 3e-2, errs


def _scene_geometry(tile_seed, G, V, hw, cell):
    from oracle import bev_mapper as obm, grids as ogrids, streetview_encoder as osv
    from snap_b200 import synthetic
    data = synthetic.make_tile(tile_seed, V, hw, G, spacing=0.5, same_side=True)
    ocam, oT = to_oracle_geometry(data, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz, _ = obm.build_xyz_query(ogrids.Grid2D((G, G), cell), oT.t)
    p2d, vis, depth, _ = osv.project_points_to_views(oT, ocam, xyz.reshape(-1, 3))
    return xyz.shape[2], p2d, vis, depth


def test_frozen_encoder_step_backward_matches_autograd():
    """`localizer_step.FrozenEncoderBackward`: map side (street-view plane fused with an aerial plane by the modality max)
    AND query side (its own lift; 60 points = a row count the split-K kernels need padded) through ONE shared BEV mapper,
    against one autograd graph with shared parameter leaves -- the gradients of both sides add up."""
    from snap_b200 import configs, localizer_step, params, pose_estimation, streetview_encoder as sve
    hw, hf, wf, D, cell = (64, 96), 16, 24, 32, 0.2
    rng = np.random.default_rng(29)
    Gm, Vm = 16, 3
    Zm, p2d_m, vis_m, depth_m = _scene_geometry(6, Gm, Vm, hw, cell)
    Xq, Yq, Vq = 10, 6, 2                                     # 60 query "points" (columns); 60 % 16 != 0
    from oracle import bev_mapper as obm, grids as ogrids, streetview_encoder as osv
    from snap_b200 import synthetic
    dq = synthetic.make_tile(9, Vq, hw, 16, spacing=0.5, same_side=True)
    ocam, oT = to_oracle_geometry(dq, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz_q, _ = obm.build_xyz_query(ogrids.Grid2D((16, 16), cell), oT.t)
    seen = osv.project_points_to_views(oT, ocam, xyz_q.reshape(-1, 3))[1].any(-1).reshape(xyz_q.shape[:3]).sum(-1)
    best = max(((seen[i:i + Xq, j:j + Yq].sum(), i, j) for i in range(16 - Xq + 1) for j in range(16 - Yq + 1)))
    xyz_q = xyz_q[best[1]:best[1] + Xq, best[2]:best[2] + Yq]   # the best-seen block of columns stands in for the FoV points
    Zq = xyz_q.shape[2]
    p2d_q, vis_q, depth_q, _ = osv.project_points_to_views(oT, ocam, xyz_q.reshape(-1, 3))
    Nq, cells_m = Xq * Yq, Gm * Gm
    assert len(p2d_q) != len(p2d_m) and vis_q.any(-1).mean() > 0.02
    svp = params.round_to_bf16(params.perturb_affine(rng, {"proj_mlp": params.init_mlp(rng, 128, (160,)),
                                                           "fusion_mlp": params.init_mlp(rng, 257, (256, 128))}))
    mp = {"kernel": bf16_np(rng.standard_normal((128, D)) * 0.2), "bias": bf16_np(rng.standard_normal(D) * 0.05)}
    enc_m = bf16_np(rng.standard_normal((Vm * hf * wf, 128)))
    enc_q = bf16_np(rng.standard_normal((Vq * hf * wf, 128)))
    aerial = bf16_np(rng.standard_normal((cells_m, 128)) * 0.05)
    P1, temp = 40, 0.1
    q_xy = np.stack([rng.uniform(0.2, 1.5, Nq), rng.uniform(-0.8, 0.8, Nq)], -1).astype(F)
    poses = np.stack([rng.uniform(-0.5, 0.5, P1), rng.uniform(0.2, 1.5, P1), rng.uniform(0.8, 2.4, P1)], -1).astype(F)

    # ---- reference: one autograd graph with shared leaves ------------------------------------------------------------
    tp, x_m, tm = chain_forward(svp, enc_m, p2d_m, vis_m, depth_m, Vm, hf, wf, cells_m, Zm, rd_bf16)
    _, x_q, tq = chain_forward(svp, enc_q, p2d_q, vis_q, depth_q, Vq, hf, wf, Nq, Zq, rd_bf16, tp=tp)
    K, bias = torch.from_numpy(mp["kernel"]).requires_grad_(True), torch.from_numpy(mp["bias"]).requires_grad_(True)
    T = torch.tensor(temp, requires_grad=True)

    def head(plane, valid):
        y = rd_bf16(rd_bf16(plane @ K) + bias)
        nrm = y.norm(dim=-1, keepdim=True)
        return torch.where(valid[:, None] & (nrm >= 1e-5), rd_bf16(y / nrm.clamp(min=1e-30)), torch.zeros(()))
    a_t = torch.from_numpy(aerial)
    sv_masked = torch.where(tm["plane_valid"][:, None], tm["plane"], torch.full((), -float("inf")))
    fused = rd_bf16(torch.stack([sv_masked, a_t], 1).amax(1))                                  # bev_mapper.py:225-252
    fm = head(fused, torch.ones(cells_m, dtype=torch.bool))
    fq = head(tq["plane"], tq["plane_valid"])
    sim = torch.relu(rd_bf16(fq @ fm.T))
    valid_q = tq["plane_valid"].numpy()
    w = 1.0 / max(int(valid_q.sum()), 1)
    sc = pose_scores((sim * torch.exp(T) * w).reshape(Nq, Gm, Gm), pose_uv(poses, q_xy, cell), valid_q, None, False)
    nll(sc, np.zeros(P1, bool)).backward()

    # ---- product: the chained plans -------------------------------------------------------------------------------------
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    nump = lambda a: a.detach().numpy()
    cfg = configs.streetview_encoder()
    lp_m = sve.fill_lift_params(cfg, Vm, hf, wf, Gm, Gm, Zm, 288)
    lp_q = sve.fill_lift_params(cfg, Vq, hf, wf, Nq, 1, Zq, 288, True)
    u8 = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.uint8))
    ctx_m = localizer_step.SceneContext(lp_m, None, bf(nump(tm["fimg"])), bf(nump(tm["crop"])), None, None, None,
                                        bf(nump(tm["plane"])), u8(nump(tm["plane_valid"])), aerial_plane=bf(aerial),
                                        fused_plane=bf(nump(fused)))
    ctx_q = localizer_step.SceneContext(lp_q, None, bf(nump(tq["fimg"])), bf(nump(tq["crop"])), None, None, None,
                                        bf(nump(tq["plane"])), u8(valid_q))
    scale = float(np.exp(F(temp)))
    maps = pose_estimation.SimilarityMaps(sim=bf(nump(sim))[None], scale=scale,
                                          point_scale=torch.from_numpy(np.where(valid_q, scale * w, 0).astype(F))[None],
                                          row_cdf=None, row_max=None, chunk_sum=None, row_sum=None, H=Gm, W=Gm)
    emu = make_lift_emulation(p2d_m, vis_m, depth_m, more={len(p2d_q): (p2d_q, vis_q, depth_q)})
    with emulated_ops(emu):
        step = localizer_step.FrozenEncoderBackward({"streetview_encoder": svp, "matching_proj": mp}, torch.device("cpu"))
        out = step.backward(maps, bf(nump(fq))[None], bf(nump(fm)).reshape(1, Gm, Gm, D), torch.from_numpy(q_xy), None,
                            torch.from_numpy(poses)[None], sc.detach()[None], cell, False, True, None, None, None,
                            [ctx_m], [ctx_q])
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    got = out["bev_mapper"]
    errs = {"matching_proj/kernel": rel(got["matching_proj"]["kernel"], K.grad.numpy()),
            "matching_proj/bias": rel(got["matching_proj"]["bias"], bias.grad.numpy()),
            "temperature": abs(float(out["temperature"]) - float(T.grad)) / (abs(float(T.grad)) + 1e-30),
            "encoder features (map)": rel(out["encoder_cotangents"]["map"][0][: Vm * hf * wf].float().numpy(), x_m.grad.numpy()),
            "encoder features (query)": rel(out["encoder_cotangents"]["query"][0][: Vq * hf * wf].float().numpy(), x_q.grad.numpy())}
    for k in ("proj_mlp", "fusion_mlp"):
        for n, d in tp[k].items():
            for a, leaf in d.items():
                errs[f"{k}/{n}/{a}"] = rel(got["streetview_encoder"][k][n][a], leaf.grad.numpy())
    print({k: round(float(v), 4) for k, v in errs.items()})
    assert float(x_q.grad.norm()) > 1e-6 and float(x_m.grad.norm()) > 1e-6, "both sides receive gradient"
    assert max(errs.values()) < 3e-2, errs


def test_plane_cotangent_to_every_encoder_weight():
    """One scene end to end: images -> `TrunkTrainer` (whole image encoder) -> proj MLP -> lift -> fusion MLP -> vertical
    max, then `LiftBackward.scene_backward` hands the cotangent of the encoder features to `TrunkTrainer.backward`.
    Against ONE autograd graph of the oracle chain: the lift / MLP gradients tightly, the encoder's arrays directionally
    (the free-running bf16 trunk is chaotic, tests/test_fpn_backward_plan_cpu.py)."""
    from oracle import image_encoder as oie, resnet as ores
    from snap_b200 import configs, encoder_train, ops, params, streetview_encoder as sve, streetview_train
    G, V, hw, cell = 16, 2, (128, 128), 0.2
    hf, wf = hw[0] // 4, hw[1] // 4
    rng = np.random.default_rng(91)
    Z, p2d, vis, depth = _scene_geometry(6, G, V, hw, cell)
    cells = G * G
    assert vis.any(-1).mean() > 0.02
    ln = lambda *s: (rng.standard_normal(s) / np.sqrt(np.prod(s[:-1]))).astype(F)
    gnp = lambda c: {"scale": (1 + 0.2 * rng.standard_normal((1, 1, 1, c))).astype(F), "bias": (0.1 * rng.standard_normal((1, 1, 1, c))).astype(F)}
    unit = lambda cin, nmid, nout: {"gn1": gnp(cin), "gn2": gnp(nmid), "gn3": gnp(nmid), "conv1": {"kernel": ln(1, 1, cin, nmid)},
                                    "conv2": {"kernel": ln(3, 3, nmid, nmid)}, "conv3": {"kernel": ln(1, 1, nmid, nout)},
                                    "conv_proj": {"kernel": ln(1, 1, cin, nout)}}
    enc = {"root_block": {"conv_root": {"kernel": ln(7, 7, 3, 64)}}, "block1": {"unit01": unit(64, 64, 256)},
           "block2": {"unit01": unit(256, 128, 512)}, "block3": {"unit01": unit(512, 256, 1024)},
           "block4": {"unit01": unit(1024, 512, 2048)}}
    dec = {}
    for level, c in enumerate((2048, 1024, 512, 256)):
        dec[f"{level}_skip_norm"] = gnp(c)
        dec[f"{level}_skip_conv"] = {"kernel": ln(1, 1, c, 128)}
    pe = params.round_to_bf16({"encoder": enc, "decoder": dec})
    svp = params.round_to_bf16(params.perturb_affine(rng, {"proj_mlp": params.init_mlp(rng, 128, (160,)),
                                                           "fusion_mlp": params.init_mlp(rng, 257, (256, 128))}))
    img = rng.random((V, hw[0], hw[1], 3)).astype(F)
    dplane = bf16_np(rng.standard_normal((cells, 128)) * 0.1)
    # reference: one graph
    tt = lambda t: {k: (tt(v) if isinstance(v, dict) else torch.from_numpy(v).requires_grad_(True)) for k, v in t.items()}
    tpe = tt(pe)
    stages = ores.resnet_v2(torch.from_numpy(img), tpe["encoder"], False, rd_bf16)
    fin_ref = oie.fpn_decoder(stages[::-1], tpe["decoder"], rd_bf16)[-1]                        # [V, hf, wf, 128]
    tp, _, t = chain_forward(svp, None, p2d, vis, depth, V, hf, wf, cells, Z, rd_bf16, x=fin_ref.reshape(V * hf * wf, 128))
    (t["plane"] * torch.from_numpy(dplane)).sum().backward()
    # product plans
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    lp = sve.fill_lift_params(configs.streetview_encoder(), V, hf, wf, G, G, Z, 288)
    with emulated_ops(make_lift_emulation(p2d, vis, depth)):
        tr = encoder_train.TrunkTrainer(pe, V, hw[0], hw[1], torch.device("cpu"))
        fin = tr.forward(torch.from_numpy(img))
        rows_img = V * hf * wf
        crop = torch.relu(fin[:rows_img]).contiguous()                                        # ops.crop_relu on full-size maps
        fimg = torch.zeros((rows_img, 160), dtype=torch.bfloat16)
        lb = streetview_train.LiftBackward(svp, torch.device("cpu"))
        ops.gemm(crop, lb.W["proj_mlp/Dense_0/kernel"].t().contiguous().to(torch.bfloat16), fimg, m_rows=rows_img,
                 bias=lb.b["proj_mlp/Dense_0/bias"])
        lb.zero_grads()
        dcrop = lb.scene_backward(lp, None, fimg, crop, None, None, None, None, None, bf(dplane))
        tr.backward(dcrop[:rows_img].contiguous())
        g_lift, g_enc = lb.grads_tree(), tr.grads_tree()
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    assert rel(fin[:rows_img].float().numpy(), fin_ref.detach().numpy().reshape(rows_img, 128)) < 3e-2
    cos = {}

    def walk(gt, rt, pre):
        for k, v in rt.items():
            if isinstance(v, dict):
                walk(gt[k], v, pre + (k,))
            else:
                g, r = gt[k].reshape(-1).astype(np.float64), v.grad.numpy().reshape(-1).astype(np.float64)
                cos["/".join(pre + (k,))] = float(g @ r / (np.linalg.norm(g) * np.linalg.norm(r) + 1e-30))
    walk(g_enc, tpe, ())
    walk(g_lift, tp, ())
    low = sorted(cos.items(), key=lambda kv: kv[1])[:3]
    print(len(cos), "arrays from the plane cotangent; lowest cosines:", [(k, round(v, 4)) for k, v in low])
    assert min(cos.values()) > 0.95, low
    for k in ("fusion_mlp/Dense_1/kernel", "fusion_mlp/Dense_0/kernel", "proj_mlp/Dense_0/kernel"):
        assert cos[k] > 0.98, (k, cos[k])      # their inputs (the encoder features) already differ by ~1 % from the oracle's
