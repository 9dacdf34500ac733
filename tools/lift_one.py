"""The fused camera->BEV lift alone at the cfg2 shape (4 views 120x160x160 feature maps -> 128x128x60 voxels), B scenes:
CUDA-event time per tile of the second-generation batched kernel (one launch) and of the first-generation per-scene
kernel, the roofline fraction (algorithmic FLOPs of SURVEY §8d / measured sustained bf16 peak) and the per-role cycle
shares the kernel reports.  Also the entry point for `ncu` captures of the kernel:

    ncu --set full --clock-control none --import-source on -k regex:lift_fused2 --profile-from-start off -c 1 \
        -o gpurun_out/lift2 python tools/lift_one.py --batch 8 --reps 1

`--dense`: cameras packed 0.5 m apart looking to the same side (most visible voxels are seen by several views).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snap_b200 import bev_mapper, configs, ops, params, streetview_encoder as sve, synthetic, types  # noqa: E402
from snap_b200.image_encoder import _WeightBank  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--dense", action="store_true")
ap.add_argument("--kernels", default="v2,v1")
args = ap.parse_args()
G, V, hw_img, B = args.grid, 4, (480, 640), args.batch
hf, wf = 120, 160
dev = torch.device("cuda", 0)
F = np.float32
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F))
rng = np.random.default_rng(5)
cfg = configs.streetview_encoder()
fp = params.round_to_bf16(params.init_mlp(rng, 257, (256, 128)))
bank = _WeightBank(dev)
w0 = bank.add(fp["Dense_0"]["kernel"], False, 32)
w1 = bank.add(fp["Dense_1"]["kernel"], False)
bank.finalize(); bank.run()
b1, b2 = t(fp["Dense_0"]["bias"]).to(dev), t(fp["Dense_1"]["bias"]).to(dev)
w256 = t(fp["Dense_0"]["kernel"][256]).to(dev)
layout = dict(spacing=0.5, same_side=True) if args.dense else {}
mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
views, zs = [], []
for b in range(B):
    data = synthetic.make_tile(100 + b, V, hw_img, G, **layout)
    xs, ys, z = mapper.build_xyz_grid(data)
    views.append(torch.from_numpy(sve.pack_views(data["camera"], data["T_view2scene"], 0, (4.0, 4.0))))
    zs.append(t(z[0]))
Z = zs[0].shape[0]
views, zs = torch.stack(views).to(dev), torch.stack(zs).to(dev)
lp = sve.fill_lift_params(cfg, V, hf, wf, G, G, Z, 288)
gen = torch.Generator().manual_seed(1)
fimg = torch.randn((B, V * hf * wf, 160), generator=gen).to(torch.bfloat16).to(dev)
xs_d, ys_d = t(xs).to(dev), t(ys).to(dev)
scratch = torch.zeros(ops.lift_fused_batched_scratch_bytes(), dtype=torch.uint8, device=dev)
counter = torch.zeros(16, dtype=torch.int32, device=dev)
plane = torch.zeros((B, G * G, 128), dtype=torch.bfloat16, device=dev)
pv = torch.zeros((B, G * G), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: the feature maps come from HBM in every repetition


def run_v2():
    ops.lift_fused_batched(lp, B, views, fimg, xs_d, ys_d, zs, bank.b_mats[w0], w256, b1, bank.b_mats[w1], b2, plane, pv,
                           counter, scratch)


def run_v1():
    for b in range(B):
        ops.lift_fused(lp, views[b], fimg[b], xs_d, ys_d, zs[b], bank.b_mats[w0], w256, b1, bank.b_mats[w1], b2, plane[b],
                       pv[b], counter, scratch)


try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    tf_sus = float(peaks["bf16_tflops_sustained"])
except Exception:
    tf_sus = 1400.0
flops_tile = 2.0 * (257 * 256 + 256 * 128) * G * G * Z
out = {"shape": f"B={B} V={V} {hf}x{wf}x160 -> {G}x{G}x{Z}", "dense": args.dense, "peak_tflops_sustained": tf_sus}
for name in args.kernels.split(","):
    fn = {"v2": run_v2, "v1": run_v1}[name]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    if name == "v2":
        torch.cuda.profiler.start()
    for _ in range(args.reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    if name == "v2":
        torch.cuda.profiler.stop()
    ms_tile = float(np.median(ms)) / B
    c = counter.cpu().tolist()
    res = {"ms_per_tile": round(ms_tile, 4), "ms_min_per_tile": round(min(ms) / B, 4),
           "frac_of_sustained_peak_algorithmic": round(flops_tile / (ms_tile * 1e-3) / 1e12 / tf_sus, 4),
           "visible_voxels_last_launch": c[2], "tiles_last_launch": c[1], "valid_cells": int(pv.sum())}
    if name == "v2":
        tot_p, tot_c = max(1, sum(c[4:7])), max(1, sum(c[7:12]))
        res["producer_share"] = dict(zip(["cull+project", "wait_buffer", "gather+pool"], [round(x / tot_p, 3) for x in c[4:7]]))
        res["consumer_share"] = dict(zip(["wait_gemm1", "epilogue1", "wait_gemm2", "epilogue2", "zmax"], [round(x / tot_c, 3) for x in c[7:12]]))
        res["visible_fraction"] = round(c[2] / (B * G * G * Z), 4)
    out[name] = res
print(json.dumps(out))
