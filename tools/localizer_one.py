"""Run the sampling localizer's matching block a few times at the config-4 per-example shape (for ncu captures).

    ncu --set full -k regex:loc_pose_scoring --profile-from-start off -c 2 -o gpurun_out/loc python tools/localizer_one.py
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snap_b200 import bev_localizer as bl, ops, pose_estimation as pe, types  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--poses", type=int, default=10_000)
ap.add_argument("--retries", type=int, default=8)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--time", action="store_true")
args = ap.parse_args()
G = args.grid
dev = torch.device("cuda", 0)
grid = types.Grid2D((G, G), 0.2)
_, _, q_xy = bl.build_query_frustum_grid(0.2, 16.0, True, 72.0)
N = q_xy.shape[0]
g = torch.Generator(device="cpu").manual_seed(7)
fq = torch.nn.functional.normalize(torch.randn((1, N, 32), generator=g), dim=-1)
fm = torch.nn.functional.normalize(torch.randn((1, G, G, 32), generator=g), dim=-1).to(torch.bfloat16).to(dev)
vq = torch.rand((1, N), generator=g) < 0.6
fq = (fq * vq[..., None]).to(torch.bfloat16).to(dev)
vq = vq.to(torch.uint8).to(dev)
q_xy_d = torch.from_numpy(np.ascontiguousarray(q_xy[:, 0])).to(dev)
gen = torch.Generator(device=dev).manual_seed(3)


def run():
    maps = pe.point_similarities(fq, vq, fm, 2.0, True, None)
    poses = pe.sample_transforms_ransac_batched(gen, maps, q_xy_d, args.poses, args.retries, grid)
    sc = pe.pose_scoring_many_batched(poses, maps, q_xy_d, None, grid, False)
    bi = torch.empty((1,), dtype=torch.int32, device=dev)
    bp = torch.empty((1, 3), dtype=torch.float32, device=dev)
    ops.argmax_rows(sc, 0, bi, poses, bp)
    return pe.grid_refinement_batched(bp, maps, q_xy_d, None, grid, False)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.reps):
    refined, vol = run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("refined", refined.cpu().numpy(), "volume", tuple(vol.shape))
if args.time:   # GPU-bound timing of the two scoring calls alone (20 back-to-back launches each)
    maps = pe.point_similarities(fq, vq, fm, 2.0, True, None)
    poses = pe.sample_transforms_ransac_batched(gen, maps, q_xy_d, args.poses, args.retries, grid)
    lattice = torch.randn((1, 68921, 3), device=dev) * torch.tensor([0.05, 2.0, 2.0], device=dev) + torch.tensor([0.3, 12.0, 12.0], device=dev)
    for name, P in (("ransac", poses), ("refine", lattice.contiguous())):
        for _ in range(3):
            pe.pose_scoring_many_batched(P, maps, q_xy_d, None, grid, False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pe.pose_scoring_many_batched(P, maps, q_xy_d, None, grid, False)
        e1.record()
        torch.cuda.synchronize()
        print(f"pose_scoring[{name}, P={P.shape[1]}] {e0.elapsed_time(e1) / 20:.4f} ms  (SNAPB200_LOC_PPT={os.environ.get('SNAPB200_LOC_PPT', 'auto')})")
