"""The launch plan of the lift backward of one scene (`snap_b200/streetview_train.py::LiftBackward`) on the CPU with
the operator layer emulated (tests/ops_emulation.py) against torch autograd of the whole chain
proj MLP -> gather / pooling -> fusion MLP -> mask -> vertical max: buffers, operand layouts, masks and launch order.
The CUDA kernels behind the operators are checked on the GPU."""
import numpy as np
import torch

from lift_torch_ref import chain_reference
from ops_emulation import emulated_ops, make_lift_emulation
from util import F, bf16_np, rd_bf16, to_oracle_geometry


def test_lift_backward_launch_plan_matches_autograd():
    from oracle import bev_mapper as obm, grids as ogrids, streetview_encoder as osv
    from snap_b200 import configs, params, streetview_encoder as sve, streetview_train, synthetic
    G, V, hw = 16, 3, (64, 96)
    hf, wf = 16, 24
    rng = np.random.default_rng(11)
    data = synthetic.make_tile(6, V, hw, G, spacing=0.5, same_side=True)
    ocam, oT = to_oracle_geometry(data, 0)
    ocam = ocam.scale(np.asarray([0.25, 0.25], dtype=F))
    xyz, _ = obm.build_xyz_query(ogrids.Grid2D((G, G), 0.2), oT.t)
    Z = xyz.shape[2]
    N, cells = G * G * Z, G * G
    p2d, vis, depth, _ = osv.project_points_to_views(oT, ocam, xyz.reshape(-1, 3))
    assert vis.any(-1).mean() > 0.02 and (vis.sum(-1) >= 2).mean() > 0.005
    cfg = configs.streetview_encoder()
    svp = params.round_to_bf16(params.perturb_affine(rng, {"proj_mlp": params.init_mlp(rng, 128, (160,)),
                                                           "fusion_mlp": params.init_mlp(rng, 257, (256, 128))}))
    enc = bf16_np(rng.standard_normal((V * hf * wf, 128)))                 # cropped finest FPN level (encoder output)
    dplane = bf16_np(rng.standard_normal((cells, 128)) * 0.1)

    # ---- reference: autograd through the whole chain (bf16 materialisation points as in the forward) -----------------
    fwd, ref, ref_x = chain_reference(svp, enc, p2d, vis, depth, V, hf, wf, cells, Z, dplane, rd_bf16)

    # ---- the product's plan on the emulated operator layer -------------------------------------------------------------
    bf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16)
    lp = sve.fill_lift_params(cfg, V, hf, wf, G, G, Z, 288)
    with emulated_ops(make_lift_emulation(p2d, vis, depth)):
        lb = streetview_train.LiftBackward(svp, torch.device("cpu"))
        lb.zero_grads()
        dcrop = lb.scene_backward(lp, None, bf(fwd["fimg"]), bf(fwd["crop"]), None, None, None, bf(fwd["vol"]),
                                  torch.from_numpy(vis.any(-1).astype(np.uint8)), bf(dplane))
        got = lb.grads_tree()
        # after the fused forward there is no volume: the backward recomputes it and must give the same gradients
        lb2 = streetview_train.LiftBackward(svp, torch.device("cpu"))
        lb2.zero_grads()
        dcrop2 = lb2.scene_backward(lp, None, bf(fwd["fimg"]), bf(fwd["crop"]), None, None, None, None, None, bf(dplane))
        got2 = lb2.grads_tree()
    for k in got:
        for n in got[k]:
            for a in got[k][n]:
                assert np.array_equal(got[k][n][a], got2[k][n][a]), (k, n, a)
    assert torch.equal(dcrop, dcrop2)
    worst = 0.0
    for k in ("proj_mlp", "fusion_mlp"):
        for n, d in ref[k].items():
            for a, r in d.items():
                g = got[k][n][a]
                err = np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
                worst = max(worst, err)
                assert g.shape == r.shape and np.linalg.norm(r) > 1e-3 and err < 3e-2, (k, n, a, err, np.linalg.norm(r))
    r = ref_x
    e_x = np.linalg.norm(dcrop[: V * hf * wf].float().numpy() - r) / np.linalg.norm(r)
    print(f"worst relative parameter-gradient error {worst:.4f}; encoder-feature cotangent {e_x:.4f}")
    assert e_x < 3e-2


def test_matching_head_and_fusion_backward_plan_matches_autograd():
    """`MatchingHeadBackward` (Dense 128 -> 32, L2 normalisation, mask; modality max in front) on the emulated operator
    layer vs torch autograd of the same head (the forward is oracle.bev_mapper.matching_head / vertical_pooling_max)."""
    from oracle import bev_mapper as obm
    from snap_b200 import streetview_train
    rng = np.random.default_rng(21)
    cells, C = 16 * 16, 128
    sv = bf16_np(np.round(rng.standard_normal((cells, C)) * 4) / 4)          # coarse values: ties with the aerial plane
    ae = bf16_np(np.round(rng.standard_normal((cells, C)) * 4) / 4)
    sv_valid = rng.random(cells) < 0.6
    sv = sv * sv_valid[:, None]
    mp = {"kernel": bf16_np(rng.standard_normal((C, 32)) * 0.1), "bias": bf16_np(rng.standard_normal(32) * 0.05)}
    dmatch = bf16_np(rng.standard_normal((cells, 32)) * 0.1)
    # reference
    a, b = torch.from_numpy(sv).requires_grad_(True), torch.from_numpy(ae).requires_grad_(True)
    K, bias = torch.from_numpy(mp["kernel"]).requires_grad_(True), torch.from_numpy(mp["bias"]).requires_grad_(True)
    va = torch.from_numpy(sv_valid)[:, None]
    stacked = torch.stack([torch.where(va, a, torch.full((), -float("inf"))), b], 1)       # bev_mapper.py:247-252
    fused = stacked.amax(1)
    fused_bf = rd_bf16(fused)
    y = rd_bf16(rd_bf16(fused_bf @ K) + bias)
    z = y / y.norm(dim=-1, keepdim=True)                                                     # all norms >> eps here
    (rd_bf16(z) * torch.from_numpy(dmatch)).sum().backward()
    ref_fwd = obm.matching_head(fused.detach().numpy(), np.ones(cells, bool), mp, rd_bf16)
    assert np.abs(ref_fwd - rd_bf16(z).detach().numpy()).max() < 1e-2                      # same head as the oracle's
    # plan
    bf = lambda t: torch.from_numpy(np.ascontiguousarray(t, dtype=F)).to(torch.bfloat16)
    with emulated_ops():
        mh = streetview_train.MatchingHeadBackward(mp, torch.device("cpu"))
        dplane = mh.backward(bf(fused.detach().numpy()), torch.ones(cells, dtype=torch.uint8), bf(dmatch))
        da, db = mh.fusion_backward(bf(sv), torch.from_numpy(sv_valid.astype(np.uint8)), bf(ae), dplane[:cells].contiguous())
    rel = lambda g, r: np.linalg.norm(g - r) / (np.linalg.norm(r) + 1e-30)
    e = dict(kernel=rel(mh.g["kernel"].numpy(), K.grad.numpy()), bias=rel(mh.g["bias"].numpy(), bias.grad.numpy()),
             sv=rel(da.float().numpy(), a.grad.numpy()), aerial=rel(db.float().numpy(), b.grad.numpy()))
    print({k: round(float(v), 5) for k, v in e.items()})
    assert (stacked[:, 0] == stacked[:, 1]).float().mean() > 0.01, "the fixture plants ties"
    assert all(v < 2e-2 for v in e.values()), e
    assert min(float(t.grad.norm()) for t in (a, b, K, bias)) > 1e-3
    assert not da.float().numpy()[~sv_valid].any()
