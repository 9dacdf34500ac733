"""Default hyper-parameters of the hot path, mirroring `snap/configs/defaults.py` field for field
(ml_collections is not installed, so a small attribute-dict stands in for ConfigDict)."""
from __future__ import annotations

import copy
from typing import Any, Iterable


class ConfigDict(dict):
    """dict with attribute access and `.get`, the subset of ml_collections.ConfigDict the modules use."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def mlp() -> ConfigDict:  # defaults.py:122-129
    return ConfigDict(activation="relu", layers=None, apply_input_activation=False)


def resnet(name: str = "R50") -> ConfigDict:  # defaults.py:132-170
    ret = ConfigDict(width=1, depth=50, limit_num_blocks=4, skip_root_block=False,
                     checkpoint_blocks=False, checkpoint_units=False, pretrained_path=None)
    if name == "R50":
        pass
    elif name == "R101":
        ret.update(depth=101, checkpoint_blocks=True, checkpoint_units=True)
    elif name == "R152x2":
        ret.update(width=2, depth=152, limit_num_blocks=3, checkpoint_blocks=True, checkpoint_units=True)
    else:
        raise ValueError(f"Unknown ResNet name: {name}")
    return ret


def image_encoder() -> ConfigDict:  # defaults.py:173-182
    return ConfigDict(encoder_name="resnet", encoder=resnet(), output_dim=128, num_pyr_levels=None)


def aerial_encoder() -> ConfigDict:  # defaults.py:185-188
    enc = image_encoder()
    enc.encoder.skip_root_block = True
    return enc


def streetview_encoder() -> ConfigDict:  # defaults.py:201-223
    feature_dim = 128
    fusion = mlp()
    fusion.layers = (feature_dim * 2, feature_dim)
    proj = mlp()
    proj.apply_input_activation = True
    return ConfigDict(image_encoder=image_encoder(), feature_dim=feature_dim, fusion=fusion, proj_mlp=proj,
                      depth_mlp=None, do_weighted_fusion=True, num_scale_bins=32, top_k_view_selection=4,
                      depth_min_max=(1.0, 32.0), fusion_add_minmax=False, fusion_use_variance=True,
                      max_view_distance=None, pretrained_path=None)


def vertical_pooling() -> ConfigDict:  # defaults.py:226-235
    fusion = mlp()
    fusion.layers = (256, 128)
    return ConfigDict(pooling="max", mlp=fusion)


def bev_mapper(modalities: Iterable[str] = ("streetview", "aerial")) -> ConfigDict:  # defaults.py:238-270
    cfg = ConfigDict(streetview_encoder=None, scene_z_offset=4.0, scene_z_offset_range=(-2, 2),
                     scene_z_height=12.0, pooling=vertical_pooling(), aerial_encoder=None,
                     semantic_encoder=None, modality_fusion=vertical_pooling(), bev_net=None,
                     matching_dim=32, normalize_matching_features=True, add_confidence=False,
                     apply_modality_dropout=True, pretrained_path=None)
    for m in modalities:
        if m == "streetview":
            cfg.streetview_encoder = streetview_encoder()
        elif m == "aerial":
            cfg.aerial_encoder = aerial_encoder()
        else:
            raise ValueError(f"Unknown / unsupported modality: {m}")
    return cfg


def semantic_net() -> ConfigDict:  # defaults.py:286-335 with the head of configs/train_semantics.py:27-30
    return ConfigDict(bev_mapper=bev_mapper(), decoder_type="resnet_stage", decoder_dim=256, mlp_num_layers=2,
                      resnet_num_units=2, apply_random_flip=False,
                      area_classes=("crosswalk", "sidewalk", "road", "terrain", "building"),
                      object_classes_exclusive=("fence", "pole", "tree"),
                      object_classes_independent=("traffic_sign", "traffic_light", "street_light"))


def get_block_desc(depth):  # snap/models/resnet.py:158-167
    if isinstance(depth, list):
        depth = tuple(depth)
    return {26: [2, 2, 2, 2], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3],
            200: [3, 24, 36, 3]}.get(depth, depth)


def bev_localizer() -> ConfigDict:  # defaults.py:343-361
    return ConfigDict(bev_mapper=bev_mapper(), bev_mapper_query=None, add_confidence_query=False,
                      add_confidence_map=False, mask_score_out_of_bounds=False, clip_negative_scores=True,
                      add_temperature=True, init_temperature=2.0, num_pose_samples=None,
                      num_pose_sampling_retries=1, query_frustum_depth=16.0, filter_points_in_fov=False,
                      threshold_remove_accurate_poses=None, do_grid_refinement=False)
