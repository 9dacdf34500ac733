"""BASELINE.json configs[4] ("semantic-mapping fine-tune head on frozen BEV features") per GPU: B scenes per step =
frozen BEVMapper forward (4 StreetView views 640x480 -> 128 x 128 BEV) + one training step of the semantic head
('mlp' decoder, or `--decoder resnet_stage` = the decoder snap/configs/train_semantics.py:27-30 fine-tunes: forward, loss,
backward, gradient mean over ranks, Adam).  CUDA-event timing after warm-up.

    python tools/bench_cfg5.py [--batch 4] [--steps 10] [--decoder mlp|resnet_stage]    # one GPU
    python -m torch.distributed.run --nproc-per-node N tools/bench_cfg5.py    # N GPUs (NCCL all-reduce of the head grads)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from snap_b200 import bev_mapper, configs, parallel, params, semantic_net, synthetic, types  # noqa: E402

F = np.float32
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--decoder", choices=("mlp", "resnet_stage"), default="mlp")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
G, hw, B = 128, (480, 640), args.batch
GT = ("road", "crosswalk", "sidewalk", "terrain", "building", "fence", "pole", "tree", "traffic_sign", "traffic_light", "street_light")
rng = np.random.default_rng(5)
cfg = configs.semantic_net()
if args.decoder == "mlp":
    cfg.decoder_type, cfg.decoder_dim, cfg.mlp_num_layers = "mlp", 128, 2
cfg.bev_mapper = configs.bev_mapper(("streetview",))
cfg.area_frequencies = tuple(zip(cfg.area_classes, (0.036434, 0.226553, 0.446990, 0.085374, 0.204649)))
cfg.object_frequencies = (("fence", 0.006257), ("pole", 0.001172), ("tree", 0.001924), ("traffic_sign", 0.000960),
                          ("traffic_light", 0.000559), ("street_light", 0.000738), ("void", 0.988391))
grid = types.Grid2D((G, G), 0.2)
mapper = bev_mapper.BEVMapper(cfg.bev_mapper, grid)
mp = params.round_to_bf16(params.init_bev_mapper(np.random.default_rng(7), cfg.bev_mapper))
if args.decoder == "mlp":
    hp = params.round_to_bf16(params.init_mlp(np.random.default_rng(8), 128, (128, 128, 12)))
else:
    hp = params.round_to_bf16(params.init_semantic_decoder(np.random.default_rng(8), cfg))
data = synthetic.make_tile(rank * 100 + 3, 4, hw, G, batch=B)
data["images"] = torch.from_numpy(data["images"]).to(dev)
masks = np.random.default_rng(9 + rank).random((B, G, G, len(GT))) < 0.2
model = semantic_net.SemanticNetModel(cfg, GT)
if args.decoder == "mlp":
    trainer = semantic_net.MLPHeadTrainer(cfg, hp, dev, lr=5e-5)
else:
    from snap_b200 import semantic_train
    trainer = semantic_train.StageHeadTrainer(cfg, hp, dev, lr=5e-5)
labels = {"rasters": {"gt_semantics": torch.from_numpy(masks.view(np.uint8)).to(dev)}}   # resident, like the images


def step():
    plane = mapper.apply({"params": mp}, dict(data))["bev_features"]      # frozen (train_semantics.py:35-36)
    return trainer.train_step(plane, model, labels)[0]


def head_only(plane):
    return trainer.train_step(plane, model, labels)[0]


def timed(fn, steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return parallel.max_over_ranks(e0.elapsed_time(e1) / steps, dev)


for _ in range(args.warmup):
    loss = step()
ms = timed(step, args.steps)
plane = mapper.apply({"params": mp}, dict(data))["bev_features"]
plane = types.FeaturePlane(plane.features.clone(), plane.valid.clone())
ms_head = timed(lambda: head_only(plane), args.steps)
if rank == 0:
    print(json.dumps({"workload": f"cfg5 per GPU: {B} scenes per step = frozen BEV forward (4 views 640x480, G=128) + head training "
                                  f"step ({'mlp decoder 128-128-128-12' if args.decoder == 'mlp' else 'resnet_stage decoder: Dense 128-256, 2 bottleneck units, MLP 256-256-12'}"
                                  f": forward, loss, backward, all-reduce, Adam); bf16",
                      "n_gpus": world, "batch_per_gpu": B, "ms_per_step": round(ms, 3), "head_train_step_ms": round(ms_head, 3),
                      "scenes_per_s": round(world * B / (ms * 1e-3), 1), "loss": float(loss.mean().item())}), flush=True)
if world > 1:
    dist.destroy_process_group()
