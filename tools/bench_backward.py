"""CUDA-event timings of the backward plans at the BASELINE configs[3] per-example sizes (one GPU, synthetic inputs):

  * `streetview_train.LiftBackward.scene_backward` for one map scene (V = 4 views, 120 x 160 feature maps, 128 x 128 x 60 voxels),
    with the forward's volume and in the volume-free form that follows the fused forward;
  * `localizer_train.LocalizerLossBackward.backward` (N = 4652 field-of-view points, 128 x 128 map, 10,001 scored poses);
  * `streetview_train.MatchingHeadBackward.backward` on a 128 x 128 plane.

    python tools/bench_backward.py [--steps 10] [--warmup 3]

Prints one JSON line per plan (ms per call, and the time of the scatter-add / pose-scoring backward kernels alone)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snap_b200 import bev_localizer, bev_mapper, configs, localizer_train, ops, params, pose_estimation  # noqa: E402
from snap_b200 import streetview_encoder as sve, streetview_train, synthetic, types  # noqa: E402
from snap_b200.image_encoder import _WeightBank  # noqa: E402

F = np.float32
ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
rng = np.random.default_rng(0)


def timed(fn):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(dev)
bf = lambda a: t(a).to(torch.bfloat16)

# ---- lift backward ---------------------------------------------------------------------------------------------------
G, V, hw, hf, wf = 128, 4, (480, 640), 120, 160
data = synthetic.make_tile(3, V, hw, G)
mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
xs, ys, zs = mapper.build_xyz_grid(data)
Z = zs.shape[1]
N, cells, rows_img = G * G * Z, G * G, V * hf * wf
cfg = configs.streetview_encoder()
svp = params.round_to_bf16({"proj_mlp": params.init_mlp(rng, 128, (160,)), "fusion_mlp": params.init_mlp(rng, 257, (256, 128))})
lp = sve.fill_lift_params(cfg, V, hf, wf, G, G, Z, 288)
views = torch.from_numpy(sve.pack_views(data["camera"], data["T_view2scene"], 0, (4.0, 4.0))).to(dev)
bank = _WeightBank(dev)
wp = bank.add(svp["proj_mlp"]["Dense_0"]["kernel"], False)
w0 = bank.add(svp["fusion_mlp"]["Dense_0"]["kernel"], False, 32)
w1 = bank.add(svp["fusion_mlp"]["Dense_1"]["kernel"], False)
bank.finalize(); bank.run()
crop = torch.relu(bf(rng.standard_normal((rows_img, 128))))
fimg = torch.zeros((rows_img, 160), dtype=torch.bfloat16, device=dev)
ops.gemm(crop, bank.b_mats[wp], fimg, m_rows=rows_img, bias=t(svp["proj_mlp"]["Dense_0"]["bias"]))
stats = torch.zeros((N, 288), dtype=torch.bfloat16, device=dev)
valid = torch.zeros(N, dtype=torch.uint8, device=dev)
xs_d, ys_d, zs_d = t(xs), t(ys), t(zs[0])
ops.lift_gather_pool(lp, views, fimg, xs_d, ys_d, zs_d, stats, valid)
hid = torch.zeros((N, 256), dtype=torch.bfloat16, device=dev)
vol = torch.zeros((N, 128), dtype=torch.bfloat16, device=dev)
ops.gemm(stats, bank.b_mats[w0], hid, m_rows=N, seg_k=288, bias=t(svp["fusion_mlp"]["Dense_0"]["bias"]), relu=True)
ops.gemm(hid, bank.b_mats[w1], vol, m_rows=N, bias=t(svp["fusion_mlp"]["Dense_1"]["bias"]), row_mask=valid)
dplane = bf(rng.standard_normal((cells, 128)) * 0.1)
lb = streetview_train.LiftBackward(svp, dev)
ms_with = timed(lambda: lb.scene_backward(lp, views, fimg, crop, xs_d, ys_d, zs_d, vol, valid, dplane))
ms_free = timed(lambda: lb.scene_backward(lp, views, fimg, crop, xs_d, ys_d, zs_d, None, None, dplane))
dstats = bf(rng.standard_normal((N, 288)) * 0.1)
gimg = torch.zeros((V, hf, wf, 160), dtype=torch.float32, device=dev)
ms_scatter = timed(lambda: ops.lift_gather_pool_backward(lp, views, fimg, xs_d, ys_d, zs_d, dstats, gimg))
print(json.dumps({"plan": "LiftBackward.scene_backward", "workload": f"V={V}, {hf}x{wf}x160 maps, {G}x{G}x{Z} voxels",
                  "visible_voxels": int(valid.sum()), "ms_with_volume": round(ms_with, 3), "ms_volume_free": round(ms_free, 3),
                  "ms_scatter_add_kernel": round(ms_scatter, 3)}), flush=True)

# ---- matching head backward ------------------------------------------------------------------------------------------------
mp = {"kernel": (rng.standard_normal((128, 32)) * 0.1).astype(F), "bias": np.zeros(32, F)}
mh = streetview_train.MatchingHeadBackward(mp, dev)
plane = bf(rng.standard_normal((cells, 128)))
pvalid = torch.ones(cells, dtype=torch.uint8, device=dev)
dmatch = bf(rng.standard_normal((cells, 32)) * 0.1)
print(json.dumps({"plan": "MatchingHeadBackward.backward", "workload": f"{G}x{G} plane",
                  "ms": round(timed(lambda: mh.backward(plane, pvalid, dmatch)), 3)}), flush=True)

# ---- localizer loss backward -------------------------------------------------------------------------------------------------
_, _, q = bev_localizer.build_query_frustum_grid(0.2, 16.0, True, 72.0)
q_xy = np.ascontiguousarray(q[:, 0])
Nq, D, P1, B, cell = len(q_xy), 32, 10001, 1, 0.2
fq = bf(rng.standard_normal((B, Nq, D)) / np.sqrt(D))
fm = bf(rng.standard_normal((B, G, G, D)) / np.sqrt(D))
vq = torch.from_numpy((rng.random((B, Nq)) < 0.6).astype(np.uint8)).to(dev)
maps = pose_estimation.point_similarities(fq, vq, fm, 0.0, True)
poses = t(np.stack([rng.uniform(-3.14, 3.14, (B, P1)), rng.uniform(2.0, 22.0, (B, P1)), rng.uniform(4.0, 22.0, (B, P1))], -1))
vj = torch.ones((B, G, G), dtype=torch.uint8, device=dev)
scores = pose_estimation.pose_scoring_many_batched(poses, maps, t(q_xy), vj, types.Grid2D((G, G), cell), True)
loc = localizer_train.LocalizerLossBackward(dev)
ms_loc = timed(lambda: loc.backward(maps, fq, fm, t(q_xy), vj, poses, scores, cell, True, True))
dscores = t(rng.standard_normal((B, P1)) * 1e-3)
dsim = torch.zeros((B, Nq + 12, G * G), dtype=torch.bfloat16, device=dev)
ms_ps = timed(lambda: ops.loc_pose_scoring_backward(maps.sim, maps.point_scale, t(q_xy), vj, poses, dscores, G, G, cell, True, True, dsim))
print(json.dumps({"plan": "LocalizerLossBackward.backward", "workload": f"N={Nq} points, {G}x{G} map, P={P1} poses, B={B}",
                  "valid_points": int(vq.sum()), "ms": round(ms_loc, 3), "ms_pose_scoring_backward_kernel": round(ms_ps, 3)}), flush=True)

# ---- whole image encoder: training forward + backward (R50 + FPN, one tile = 4 views padded to 512 x 672) ------------------
from snap_b200 import encoder_train  # noqa: E402

ep = params.round_to_bf16(params.init_image_encoder(np.random.default_rng(3), configs.image_encoder()))
Hp, Wp, nv = 512, 672, 4
tr = encoder_train.TrunkTrainer(ep, nv, Hp, Wp, dev)
img = torch.rand((nv, Hp, Wp, 3), dtype=torch.float32, device=dev)
tr.forward(img)
dfin = bf(rng.standard_normal((nv * (Hp // 4) * (Wp // 4), 128)) * 0.01)
ms_ef = timed(lambda: tr.forward(img))
ms_eb = timed(lambda: tr.backward(dfin))
print(json.dumps({"plan": "TrunkTrainer forward / backward", "workload": f"R50 + FPN, {nv} views {Hp}x{Wp}",
                  "ms_training_forward": round(ms_ef, 3), "ms_backward": round(ms_eb, 3)}), flush=True)
