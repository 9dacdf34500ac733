"""BASELINE.json configs[3] ("full localization: BEV + (x,y,theta) correlation, 36 rotations") per GPU, forward only:
B examples per step, each = map tile (4 StreetView views + aerial raster, G = 128) + query BEV (one view) + pose search.
Two pose searches are timed: the exhaustive voting of `pose_exhaustive_voting.py` (query lifted on the map grid) and
the sampling localizer of `bev_localizer.py` (query lifted on the 4,652 field-of-view points, 10,000 poses x 8 retries,
41^3 refinement).  CUDA-event timing after warm-up, eager launches and one captured CUDA graph per pipeline.

    python tools/bench_cfg4.py [--batch 4] [--steps 10]      # prints one JSON line, also used for profiles/
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snap_b200 import bev_localizer, bev_mapper, configs, params, pose_exhaustive_voting as pv, synthetic, types  # noqa: E402

F = np.float32
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
G, R, hw, B = 128, 36, (480, 640), args.batch
dev = torch.device("cuda", 0)
grid = types.Grid2D((G, G), 0.2)
rng = np.random.default_rng(11)
cfg = configs.bev_localizer()
cfg.bev_mapper = configs.bev_mapper(("streetview", "aerial"))
cfg.filter_points_in_fov, cfg.num_pose_samples, cfg.num_pose_sampling_retries, cfg.do_grid_refinement = True, 10_000, 8, True
loc = bev_localizer.BEVLocalizer(cfg, None, grid)
mp = params.round_to_bf16(params.init_bev_mapper(rng, cfg.bev_mapper))
variables = {"params": loc.init_params(mp)}
data = synthetic.make_tile(71, 4, hw, G, aerial=True, batch=B)
T, cam = data["T_view2scene"], data["camera"]
z_off = (np.median(T.t[..., -1].astype(F), axis=-1).astype(F) - F(4.0)).astype(F)
data["z_offset"] = z_off
data["images"] = torch.from_numpy(data["images"]).to(dev)
data["rasters"] = {"rgb": torch.from_numpy(data["rasters"]["rgb"]).to(dev)}
data["staging_slot"] = 300          # explicit staging slots: map and query calls share one captured graph
v = 2
query_same_frame = {"images": data["images"][:, [v]].contiguous(),
                    "camera": types.Camera(wh=cam.wh[:, [v]].copy(), f=cam.f[:, [v]].copy(), c=cam.c[:, [v]].copy()),
                    "T_view2scene": types.Transform3D(R=T.R[:, [v]].copy(), t=T.t[:, [v]].copy()), "z_offset": z_off,
                    "staging_slot": 301}
shift = np.concatenate([np.round(T.t[:, v, :2] / 0.2) * 0.2, np.zeros((B, 1))], -1).astype(F)
query_own_frame = dict(query_same_frame)
query_own_frame["T_view2scene"] = types.Transform3D(R=T.R[:, [v]].copy(), t=(T.t[:, [v]] - shift[:, None]).astype(F))
T_q2m = types.Transform3D(R=np.tile(np.eye(3, dtype=F), (B, 1, 1)), t=shift)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
u_fix = torch.rand((B, 10_000 * 8 * 2, 2), dtype=torch.float32, device=dev, generator=gen)
mapper = loc.bev_mapper


def voting_step():
    pm = mapper.apply({"params": mp}, dict(data))["bev_matching"]
    pq = mapper.apply({"params": mp}, dict(query_same_frame), is_query=True)["bev_matching"]
    return pv.exhaustive_pose_voting(pq, pm, R, grid)


def sampling_step():
    pm = mapper.apply({"params": mp}, dict(data))["bev_matching"]
    pq = mapper.apply({"params": mp}, {**query_own_frame, "xy_bev": loc.q_xy_p}, is_query=True)["bev_matching"]
    return loc.match(variables["params"], pq, pm, None, None, None, uniforms=u_fix)["map_t_query"]


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def capture(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


out = {"workload": f"cfg4 forward per GPU: {B} examples per step = map tile (4 views 640x480 + aerial, G=128) + query BEV "
                   f"(1 view) + pose search; bf16; one B200", "batch": B, "steps": args.steps}
pipelines = (("exhaustive_voting_R36", voting_step), ("sampling_localizer_10000x8_refine", sampling_step))
for name, fn in pipelines:
    for _ in range(args.warmup):
        fn()
    ms_eager = timed(fn, args.steps)
    out[name] = {"ms_per_step_eager": round(ms_eager, 3), "ms_per_step_graph": None,
                 "examples_per_s": round(B / (ms_eager * 1e-3), 1)}
print(json.dumps({**out, "stage": "eager"}), flush=True)
for name, fn in pipelines:      # one captured CUDA graph per pipeline (an optimisation of the launch path)
    try:
        g = capture(fn)
        for _ in range(2):
            g.replay()
        ms_graph = timed(g.replay, args.steps)
        out[name]["ms_per_step_graph"] = round(ms_graph, 3)
        out[name]["examples_per_s"] = round(B / (min(ms_graph, out[name]["ms_per_step_eager"]) * 1e-3), 1)
    except Exception as e:
        out[name + "_graph_error"] = repr(e)[:200]
        break
print(json.dumps({**out, "stage": "final"}), flush=True)
