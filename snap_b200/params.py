"""Flax-compatible parameter trees (names and layouts of SURVEY.md Appendix B) with NumPy leaves,
drawn with the initialisers the reference uses.  Used for random-init weights (no checkpoints are
released, `README.md:31-33`) and as the interchange format for real checkpoints."""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import configs

F = np.float32


def _trunc_normal(rng, shape, std):
    # jax.nn.initializers.variance_scaling(..., 'truncated_normal'): N(0,1) truncated to [-2,2], rescaled
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return (x * std / 0.87962566103423978).astype(F)


def lecun_normal(rng, shape):
    """flax nn.Conv default kernel_init; shape HWIO, fan_in = kh*kw*in."""
    fan_in = int(np.prod(shape[:-1]))
    return _trunc_normal(rng, shape, np.sqrt(1.0 / fan_in))


def glorot_uniform(rng, shape):
    """`snap/models/layers.py:69`; shape [in, out]."""
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, shape).astype(F)


def _gn(c):
    return {"scale": np.ones((1, 1, 1, c), F), "bias": np.zeros((1, 1, 1, c), F)}


def init_resnet(rng, cfg) -> Dict:
    """`snap/models/resnet.py:170-216` parameter tree."""
    width = int(64 * cfg.width)
    blocks = configs.get_block_desc(cfg.depth)
    if cfg.limit_num_blocks is not None:
        blocks = blocks[: cfg.limit_num_blocks]
    p: Dict = {}
    if cfg.skip_root_block:
        p["conv_root"] = {"kernel": lecun_normal(rng, (3, 3, 3, width))}
    else:
        p["root_block"] = {"conv_root": {"kernel": lecun_normal(rng, (7, 7, 3, width))}}
    cin = width
    for i, nunits in enumerate(blocks):
        nmid = width * 2 ** i
        nout = nmid * 4
        stage = {}
        for u in range(nunits):
            unit = {"gn1": _gn(cin), "gn2": _gn(nmid), "gn3": _gn(nmid),
                    "conv1": {"kernel": lecun_normal(rng, (1, 1, cin, nmid))},
                    "conv2": {"kernel": lecun_normal(rng, (3, 3, nmid, nmid))},
                    "conv3": {"kernel": lecun_normal(rng, (1, 1, nmid, nout))}}
            if u == 0:  # cin != nout or stride != 1 (`resnet.py:121`)
                unit["conv_proj"] = {"kernel": lecun_normal(rng, (1, 1, cin, nout))}
            stage[f"unit{u + 1:02d}"] = unit
            cin = nout
        p[f"block{i + 1}"] = stage
    return p


def init_image_encoder(rng, cfg) -> Dict:
    """`snap/models/image_encoder.py:97-117` (+ FPNDecoder :42-94)."""
    enc = init_resnet(rng, cfg.encoder)
    width = int(64 * cfg.encoder.width)
    nblocks = sum(1 for k in enc if k.startswith("block"))
    levels = cfg.num_pyr_levels or nblocks
    dec = {}
    for level in range(levels):
        stage = nblocks - 1 - level  # level 0 = coarsest stage (`image_encoder.py:114`)
        cin = width * 2 ** stage * 4
        dec[f"{level}_skip_norm"] = _gn(cin)
        dec[f"{level}_skip_conv"] = {"kernel": lecun_normal(rng, (1, 1, cin, cfg.output_dim))}
    return {"encoder": enc, "decoder": dec}


def init_mlp(rng, in_dim, layers) -> Dict:
    p = {}
    for i, d in enumerate(layers):
        p[f"Dense_{i}"] = {"kernel": glorot_uniform(rng, (in_dim, d)), "bias": np.zeros((d,), F)}
        in_dim = d
    return p


def init_streetview_encoder(rng, cfg) -> Dict:
    """`snap/models/streetview_encoder.py:196-215`."""
    d, s = cfg.feature_dim, cfg.num_scale_bins
    weighted = bool(cfg.do_weighted_fusion)   # the score_max statistic exists only with weighted fusion (`:174-177`)
    stats_dim = d * (1 + int(cfg.fusion_use_variance) + 2 * int(cfg.fusion_add_minmax)) + int(weighted)
    p = {"image_encoder": init_image_encoder(rng, cfg.image_encoder)}
    if weighted:                              # `:207-213`; without it the module has no proj_mlp
        p["proj_mlp"] = init_mlp(rng, cfg.image_encoder.output_dim, (d + s,))
    elif cfg.get("depth_mlp") is not None:    # `:214-215`: per-observation MLP on [f | log10 depth | ray]
        p["depth_mlp"] = init_mlp(rng, d + 4, tuple(cfg.depth_mlp.layers))
    p["fusion_mlp"] = init_mlp(rng, stats_dim, cfg.fusion.layers)
    return p


def init_bev_mapper(rng, cfg) -> Dict:
    """`snap/models/bev_mapper.py:107-157`."""
    p: Dict = {}
    dim = None
    if cfg.streetview_encoder is not None:
        p["streetview_encoder"] = init_streetview_encoder(rng, cfg.streetview_encoder)
        dim = cfg.streetview_encoder.feature_dim
    if cfg.aerial_encoder is not None:
        p["aerial_encoder"] = init_image_encoder(rng, cfg.aerial_encoder)
        dim = cfg.aerial_encoder.output_dim
    def pooling_params(pcfg, z: int):
        """`VerticalPooling.setup` (`bev_mapper.py:48-54`): confidence_head Dense(1) or fusion_mlp over Z*C inputs."""
        if pcfg.pooling in ("weighted", "softmax"):
            return {"confidence_head": {"kernel": lecun_normal(rng, (dim, 1)), "bias": np.zeros((1,), F)}}
        if pcfg.pooling == "mlp":
            return {"fusion_mlp": init_mlp(rng, z * dim, pcfg.mlp.layers)}
        return {}

    if cfg.streetview_encoder is not None:
        z = int(round(cfg.scene_z_height / cfg.get("voxel_size", 0.2)))
        vp = pooling_params(cfg.pooling, z)
        if vp:
            p["vertical_pooling"] = vp
    if cfg.streetview_encoder is not None and cfg.aerial_encoder is not None:
        mf = pooling_params(cfg.modality_fusion, 2)
        if mf:
            p["modality_fusion"] = mf
    if cfg.add_confidence:  # nn.Sequential([nn.Dense(1)]) (`bev_mapper.py:154-157`)
        p["confidence_head"] = {"layers_0": {"kernel": lecun_normal(rng, (dim, 1)), "bias": np.zeros((1,), F)}}
    if cfg.matching_dim is not None:
        # variance_scaling(1/sqrt(matching_dim), 'fan_in', 'truncated_normal') (`bev_mapper.py:145-153`)
        std = np.sqrt((1.0 / np.sqrt(cfg.matching_dim)) / dim)
        p["matching_proj"] = {"kernel": _trunc_normal(rng, (dim, cfg.matching_dim), std),
                              "bias": np.zeros((cfg.matching_dim,), F)}
    return p


def init_semantic_decoder(rng, cfg, in_dim: int = 128) -> Dict:
    """`snap/models/semantic_net.py:145-165` (decoder_type='resnet_stage'): layers_0 Dense, layers_1 ResNetStage,
    layers_3 MLP (Appendix B names)."""
    dim = cfg.decoder_dim
    nmid = dim // 4
    k = len(cfg.area_classes) + len(cfg.object_classes_exclusive) + len(cfg.object_classes_independent) + 1
    stage = {}
    for u in range(cfg.resnet_num_units):
        stage[f"unit{u + 1:02d}"] = {"gn1": _gn(dim), "gn2": _gn(nmid), "gn3": _gn(nmid),
                                     "conv1": {"kernel": lecun_normal(rng, (1, 1, dim, nmid))},
                                     "conv2": {"kernel": lecun_normal(rng, (3, 3, nmid, nmid))},
                                     "conv3": {"kernel": lecun_normal(rng, (1, 1, nmid, dim))}}
    return {"layers_0": {"kernel": glorot_uniform(rng, (in_dim, dim)), "bias": np.zeros((dim,), F)},
            "layers_1": stage, "layers_3": init_mlp(rng, dim, (dim, k))}


def perturb_affine(rng, tree: Dict, scale: float = 0.2) -> Dict:
    """Make GroupNorm scale/bias and Dense biases non-trivial so parity tests exercise them."""
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out[k] = perturb_affine(rng, v, scale)
        elif k in ("scale", "bias"):
            out[k] = (v + scale * rng.standard_normal(v.shape)).astype(F)
        else:
            out[k] = v
    return out


def round_to_bf16(tree: Dict) -> Dict:
    """param_dtype = bf16: parameters are STORED in bf16; keep fp32 containers."""
    import torch
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out[k] = round_to_bf16(v)
        else:
            out[k] = torch.from_numpy(np.ascontiguousarray(v)).to(torch.bfloat16).float().numpy()
    return out
