"""The gradient all-reduce of the localisation model at FULL size (`snap/trainer.py:231-234`: jax.lax.pmean over the ~48 M
parameter tree, ~193 MB fp32 per step) through `parallel.GradBucket`: one in-place NCCL all-reduce (AVG) of a persistent
flat buffer.  Reports the time per call, the bus bandwidth (2 (N-1)/N x bytes / time, the NCCL convention; measured
reference for this pool: 725 GB/s at 1 GiB over 8 ranks), and the cost of the same collective when it is enqueued on a
side stream UNDER a running BEV forward (cfg2 step), i.e. how much of it the overlap hides.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_bench.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snap_b200 import bev_mapper, configs, parallel, params, synthetic, types  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

cfg = configs.bev_mapper(("streetview", "aerial"))
tree = params.init_bev_mapper(np.random.default_rng(1), cfg)
shapes = [np.asarray(v).shape for _, v in parallel._leaves(tree)]
bucket = parallel.GradBucket(shapes, dev)
nparam = sum(int(np.prod(s)) for s in shapes)
for v in bucket.views:
    v.fill_(float(rank + 1))


def timed(fn, n):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return parallel.max_over_ranks(e0.elapsed_time(e1) / n, dev)


for _ in range(3):
    bucket.allreduce_mean()
bucket.flat.fill_(float(rank + 1))
bucket.allreduce_mean()
torch.cuda.synchronize()
want = (world + 1) / 2.0
ok = bool(torch.allclose(bucket.flat[:1000], torch.full((1000,), want, device=dev)))
ms = timed(bucket.allreduce_mean, 20)

# overlap with a BEV forward: the collective goes on a side stream while the (frozen) forward of the next micro-batch runs
G, B = 128, 4
mapper = bev_mapper.BEVMapper(configs.bev_mapper(("streetview",)), types.Grid2D((G, G), 0.2))
mp = params.round_to_bf16(params.init_bev_mapper(np.random.default_rng(7), configs.bev_mapper(("streetview",))))
data = synthetic.make_tile(3 + rank, 4, (480, 640), G, batch=B)
data["images"] = torch.from_numpy(data["images"]).to(dev)
side = torch.cuda.Stream()


def step():
    mapper.apply({"params": mp}, dict(data))


def step_overlapped():
    bucket.allreduce_mean(stream=side)
    mapper.apply({"params": mp}, dict(data))
    bucket.wait()


for _ in range(3):
    step_overlapped()
ms_step = timed(step, 10)
ms_both = timed(step_overlapped, 10)
if rank == 0:
    print(json.dumps({"what": "GradBucket.allreduce_mean on the full bev_mapper parameter tree (street-view + aerial R50 encoders, heads)",
                      "n_gpus": world, "parameters": nparam, "bytes": bucket.nbytes, "leaves": len(shapes), "correct": ok,
                      "ms_per_allreduce": round(ms, 4),
                      "bus_gbs": round(2.0 * (world - 1) / world * bucket.nbytes / (ms * 1e-3) / 1e9, 1) if world > 1 else None,
                      "cfg2_forward_ms_B4": round(ms_step, 3), "forward_plus_overlapped_allreduce_ms": round(ms_both, 3),
                      "exposed_allreduce_ms": round(ms_both - ms_step, 3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
