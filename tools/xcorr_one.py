"""Run exhaustive (x, y, theta) voting a few times at the config-4 per-example shape (for ncu captures).

    ncu --set full -k regex:xcorr_rows --profile-from-start off -c 1 -o gpurun_out/xcorr python tools/xcorr_one.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snap_b200 import pose_exhaustive_voting as pv, types  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--rotations", type=int, default=36)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
G, R, B, D = args.grid, args.rotations, args.batch, 32
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(5)
fq = torch.nn.functional.normalize(torch.randn((B, G, G, D), generator=g), dim=-1).to(torch.bfloat16).to(dev)
fm = torch.nn.functional.normalize(torch.randn((B, G, G, D), generator=g), dim=-1).to(torch.bfloat16).to(dev)
vq = torch.ones((B, G, G), dtype=torch.uint8, device=dev)
vq[:, : G // 10] = 0
vm = torch.ones((B, G, G), dtype=torch.uint8, device=dev)
grid = types.Grid2D((G, G), 0.2)
for _ in range(2):
    pv.exhaustive_pose_voting(types.FeaturePlane(fq, vq), types.FeaturePlane(fm, vm), R, grid)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.reps):
    s = pv.exhaustive_pose_voting(types.FeaturePlane(fq, vq), types.FeaturePlane(fm, vm), R, grid)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("scores", tuple(s.shape), "finite", int(torch.isfinite(s).sum()))
