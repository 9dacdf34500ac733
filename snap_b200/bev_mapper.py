"""B200 implementation behind `snap.models.bev_mapper.{VerticalPooling, BEVMapper}`
(`snap/models/bev_mapper.py:40-296`)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _cache, _lib, configs, image_encoder, ops, streetview_encoder, types

F = np.float32


class VerticalPooling:
    """`bev_mapper.py:40-88`: pooling in {'max', 'sum', 'mean', 'softmax', 'weighted', 'mlp'}.

    Parameters (Flax names): 'confidence_head' {kernel [C,1], bias [1]} for 'softmax' / 'weighted',
    'fusion_mlp' {Dense_i} for 'mlp'.  The default 'max' needs none (and is normally fused into the lift kernel)."""

    POOLINGS = ("max", "sum", "mean", "softmax", "weighted", "mlp")

    def __init__(self, config=None, dtype=torch.bfloat16):
        self.config = config if config is not None else configs.vertical_pooling()
        if self.config.pooling not in self.POOLINGS:
            raise NotImplementedError(self.config.pooling)   # bev_mapper.py:53-54
        self._cache = _cache.ParamCache()

    def clear_cache(self) -> None:
        self._cache.clear()

    def _mlp_weights(self, params: Dict, device):
        def build():
            bank = image_encoder._WeightBank(device)
            f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1)).to(device)
            n = 0
            while f"Dense_{n}" in params["fusion_mlp"]:
                n += 1
            ids = [bank.add(params["fusion_mlp"][f"Dense_{i}"]["kernel"], False) for i in range(n)]
            bias = [f32(params["fusion_mlp"][f"Dense_{i}"]["bias"]) for i in range(n)]
            bank.finalize()
            return bank, ids, bias
        return self._cache.lookup(params, ("mlp", str(device)), build)

    def apply(self, variables, feature_volume: types.FeatureVolume) -> Dict:
        f, v = feature_volume.features.contiguous(), feature_volume.valid.contiguous()
        lead, Z, C = f.shape[:-2], f.shape[-2], f.shape[-1]
        cells = int(np.prod(lead))
        dev = f.device
        mode = self.config.pooling
        plane = torch.empty((*lead, C), dtype=torch.bfloat16, device=dev)
        pvalid = torch.empty(lead, dtype=torch.uint8, device=dev)
        pred: Dict = {}
        if mode == "max":
            ops.vertical_max(f, v, cells, Z, C, plane, pvalid)
        elif mode in ("sum", "mean"):
            ops.vertical_pool(mode, f, v, cells, Z, C, None, 0.0, plane, pvalid)
        elif mode in ("softmax", "weighted"):
            params = variables["params"] if "params" in variables else variables
            head = params["confidence_head"]
            w, b = self._cache.lookup(params, ("conf", str(dev)), lambda: (
                torch.from_numpy(np.ascontiguousarray(head["kernel"], dtype=F).reshape(-1)).to(dev),
                float(np.asarray(head["bias"], dtype=F).reshape(-1)[0])))
            pred["scores"] = torch.empty((*lead, Z), dtype=torch.float32, device=dev)
            pred["weights"] = torch.empty((*lead, Z), dtype=torch.float32, device=dev)
            ops.vertical_pool(mode, f, v, cells, Z, C, w, b, plane, pvalid, pred["scores"], pred["weights"])
        else:  # 'mlp' (:74-77): zero the invalid voxels, flatten each column to [Z*C], MLP, zero where no z is valid
            params = variables["params"] if "params" in variables else variables
            bank, ids, bias = self._mlp_weights(params, dev)
            bank.run()
            x = torch.zeros((max(cells, 128), Z * C), dtype=torch.bfloat16, device=dev)
            ops.mask_rows(f, v, cells * Z, C, x)
            ops.valid_any(v, cells, Z, pvalid)
            for i, wid in enumerate(ids):
                last = i == len(ids) - 1
                n_out = bank.entries[wid][2]
                y = plane.view(cells, C) if last else torch.zeros((max(cells, 128), n_out), dtype=torch.bfloat16, device=dev)
                if last and n_out != C:
                    raise NotImplementedError("the 'mlp' pooling must end in the feature dimension")
                ops.gemm(x, bank.b_mats[wid], y, m_rows=cells, bias=bias[i], relu=not last,
                         row_mask=pvalid.view(-1) if last else None)
                x = y
        pred["plane"] = types.FeaturePlane(features=plane, valid=pvalid)
        return pred

    __call__ = apply


class BEVMapper:
    """Mirror of `snap.models.bev_mapper.BEVMapper(config, grid, semantic_map_classes, dtype)`."""

    default_config = staticmethod(configs.bev_mapper)

    def __init__(self, config=None, grid: types.Grid2D = None, semantic_map_classes=None, dtype=torch.bfloat16,
                 fused_lift: bool = True):
        # fused_lift: run lift + fusion MLP + vertical pooling as one kernel (no 'feature_volume' in the
        # result); debug=True or fused_lift=False use the unfused path that materialises the volume.
        self.fused_lift = fused_lift
        self.config = config if config is not None else configs.bev_mapper()
        self.grid = grid
        self.dtype = dtype
        c = self.config
        if c.semantic_encoder is not None:
            raise NotImplementedError("semantic raster modality is out of scope (SURVEY.md §2)")
        if c.bev_net is not None:
            raise NotImplementedError("BEV network not yet implemented")  # bev_mapper.py:141-142
        if c.matching_dim is not None and not 1 <= c.matching_dim <= 256:
            raise NotImplementedError("matching_dim must be in 1..256")
        self.streetview_encoder = self.aerial_encoder = None
        if c.streetview_encoder is not None:
            self.streetview_encoder = streetview_encoder.StreetViewEncoder(c.streetview_encoder, dtype)
            self.vertical_pooling = VerticalPooling(c.pooling, dtype)
        if c.aerial_encoder is not None:
            self.aerial_encoder = image_encoder.ImageEncoder(c.aerial_encoder, dtype)
        if self.streetview_encoder is None and self.aerial_encoder is None:
            raise ValueError("Need to create at least one input encoder.")
        if self.streetview_encoder is not None and self.aerial_encoder is not None:
            self.modality_fusion = VerticalPooling(c.modality_fusion, dtype)
        self._cache: Dict = {}           # xy_bev layouts (the entry keeps the keyed array alive)
        self._wcache = _cache.ParamCache()  # device copies of the head parameters, per parameter tree

    def build_xyz_grid(self, data: Dict):
        """`bev_mapper.py:159-196` (inference): separable voxel-centre coordinates, fp32, host side."""
        t = data["T_view2scene"].t
        xy_shape = None
        xy = data.get("xy_bev")                                                             # :163
        if xy is None:
            xs, ys = self.grid.cell_centers(0), self.grid.cell_centers(1)                   # :164-165
        else:
            xy = np.asarray(xy, dtype=F)
            if xy.ndim == 4:    # batched: the kernels take one set of BEV points per call (apply() splits other batches)
                if not all(np.array_equal(xy[0], xy[b]) for b in range(1, len(xy))):
                    raise ValueError("per-example xy_bev reaches build_xyz_grid only through BEVMapper.apply")
                xy = xy[0]
            if data.get("xy_bev_cache", True):
                # a persistent point set (BEVLocalizer.q_xy_p): analysed once, and the SAME host arrays are handed out
                # so that the lift's device copy is reused (the entry keeps the key object alive, so its id is unique)
                key = ("xy_bev", id(data["xy_bev"]))
                if key not in self._cache:
                    self._cache[key] = (*self._xy_layout(xy), data["xy_bev"])
                xs, ys, xy_shape, _ = self._cache[key]
            else:
                xs, ys, xy_shape = self._xy_layout(xy)
        z_offset = data.get("z_offset")
        if z_offset is None:
            cam_h = np.median(t[..., -1].astype(F), axis=-1).astype(F)                      # :171
            z_offset = (cam_h - F(self.config.get("scene_z_offset", 4.0))).astype(F)       # :173-174
        z_offset = np.asarray(z_offset, dtype=F).reshape(-1)
        base = np.arange(0, self.config.get("scene_z_height", 12.0), self.grid.cell_size).astype(F)
        zs = ((base[None] + z_offset[:, None]).astype(F) + F(self.grid.cell_size / 2)).astype(F)  # :188-192
        if xy_shape is not None:
            data["xy_shape"] = xy_shape
        if xy is not None:
            data["xy_custom"] = True     # the lift re-uploads its (x, y) tables when these host arrays change
        return xs, ys, zs

    @staticmethod
    def _xy_layout(xy: np.ndarray):
        """[X,Y,2] BEV points -> (xs, ys, None) if they form a separable grid, else per-column (xs, ys, (X, Y))."""
        if np.array_equal(xy[..., 0], np.broadcast_to(xy[:, :1, 0], xy.shape[:2])) and \
                np.array_equal(xy[..., 1], np.broadcast_to(xy[:1, :, 1], xy.shape[:2])):
            return np.ascontiguousarray(xy[:, 0, 0]), np.ascontiguousarray(xy[0, :, 1]), None
        # arbitrary BEV points (e.g. the field-of-view filtered query frustum, [N,1,2])
        return (np.ascontiguousarray(xy[..., 0].reshape(-1)), np.ascontiguousarray(xy[..., 1].reshape(-1)),
                tuple(xy.shape[:2]))

    def split_xyz_query(self, data: Dict):
        """An explicit `data['xyz_query']` [B,X,Y,Z,3] (`bev_mapper.py:162,193-196`) in the layout the kernels take:
        the points of a voxel column share (x, y) and every column uses the same z levels per example -- which is how
        the reference itself builds the tensor (`:187-196`).  Anything else is refused."""
        xyz = np.asarray(data["xyz_query"], dtype=F)
        if xyz.ndim != 5 or xyz.shape[-1] != 3:
            raise ValueError("xyz_query must be [B,X,Y,Z,3]")
        xy, z = xyz[..., :2], xyz[..., 2]
        if not (np.array_equal(xy, np.broadcast_to(xy[:, :, :, :1], xy.shape))
                and np.array_equal(z, np.broadcast_to(z[:, :1, :1], z.shape))):
            raise NotImplementedError("xyz_query must be column-structured: (x, y) per BEV column, z levels per example")
        sub = dict(data, xy_bev=xy[:, :, :, 0], xy_bev_cache=False, z_offset=np.zeros(len(xyz), F))
        xs, ys, _ = self.build_xyz_grid(sub)
        if "xy_shape" in sub:
            data["xy_shape"] = sub["xy_shape"]
        data["xy_custom"] = True
        return xs, ys, np.ascontiguousarray(z[:, 0, 0])

    def encode_streetview(self, params, data, train, is_query, debug=False) -> Dict:
        if "xyz_query" in data and "xyz_grid" not in data:
            data = dict(data)
            data["xyz_grid"] = self.split_xyz_query(data)          # precomputed data['xyz_query'] (bev_mapper.py:162)
        if "xyz_grid" not in data:
            data = dict(data)
            data["xyz_grid"] = self.build_xyz_grid(data)
        V = data["T_view2scene"].t.shape[1]
        fused = self.fused_lift and not debug and self.config.pooling.pooling == "max" and V <= 4 \
            and not self.streetview_encoder.uses_view_selection(V) and self.streetview_encoder.default_stats \
            and not self.streetview_encoder.has_depth_mlp
        pred = self.streetview_encoder.apply({"params": params["streetview_encoder"]}, data, train, debug=debug,
                                             fused=fused)
        if not fused:
            pred["vertical_pooling"] = self.vertical_pooling.apply({"params": params.get("vertical_pooling", {})},
                                                                   pred["feature_volume"])
            pred["feature_plane"] = pred["vertical_pooling"].pop("plane")
        return pred

    def encode_aerial(self, params, aerial_rgb, train=False) -> Dict:  # bev_mapper.py:203-212
        if not isinstance(aerial_rgb, torch.Tensor):
            aerial_rgb = torch.from_numpy(np.ascontiguousarray(aerial_rgb, dtype=F)).cuda()
        pyr = self.aerial_encoder.apply({"params": params["aerial_encoder"]}, aerial_rgb, train)
        feats = pyr.features[-1]
        valid = torch.ones(feats.shape[:-1], dtype=torch.uint8, device=feats.device)
        return {"feature_plane": types.FeaturePlane(features=feats, valid=valid)}

    def fuse_neural_maps(self, planes, train: bool = False, params: Optional[Dict] = None) -> types.FeaturePlane:  # bev_mapper.py:225-252
        if not planes:
            raise ValueError("No feature plane given.")
        if len(planes) == 1:
            return planes[0]
        if self.config.modality_fusion.pooling == "max" and len(planes) == 2:
            a, b = planes
            fa, fb = a.features.contiguous(), b.features.contiguous()
            cells, C = fa.numel() // fa.shape[-1], fa.shape[-1]
            out = torch.empty_like(fa)
            vout = torch.empty(fa.shape[:-1], dtype=torch.uint8, device=fa.device)
            ops.fuse_max(fa, a.valid.contiguous(), fb, b.valid.contiguous(), cells, C, out, vout)
            return types.FeaturePlane(features=out, valid=vout)
        # other fusion modes: VerticalPooling over the stacked modality axis (:247-252)
        stacked = types.FeatureVolume(features=torch.stack([p.features for p in planes], dim=-2),
                                      valid=torch.stack([p.valid for p in planes], dim=-1))
        return self.modality_fusion.apply({"params": (params or {}).get("modality_fusion", {})}, stacked)["plane"]

    def apply(self, variables: Dict, data: Dict, train: bool = False, debug: bool = False,
              is_query: bool = False) -> Dict:
        if train:
            raise NotImplementedError("training (backward kernels) is a 'next' row of SURVEY.md §8(f)")
        params = variables["params"] if "params" in variables else variables
        xy = data.get("xy_bev")
        if xy is not None and np.ndim(xy) == 4 and len(xy) > 1 and \
                not all(np.array_equal(np.asarray(xy[0]), np.asarray(xy[b])) for b in range(1, len(xy))):
            return self._apply_per_example(variables, data, train, debug, is_query)
        pred, planes = {}, []
        if self.streetview_encoder is not None:
            pred["streetview"] = self.encode_streetview(params, data, train, is_query, debug)
            planes.append(pred["streetview"]["feature_plane"])
        if self.aerial_encoder is not None and "rasters" in data:
            pred["aerial"] = self.encode_aerial(params, data["rasters"]["rgb"], train)
            planes.append(pred["aerial"]["feature_plane"])
        if not planes:
            raise ValueError("No map encoder given.")
        plane = pred["bev_features"] = self.fuse_neural_maps(planes, train, params)
        if self.config.matching_dim is not None:  # bev_mapper.py:284-291
            dev = plane.features.device
            mp = params["matching_proj"]
            k, bvec = self._wcache.lookup(params, ("match", str(dev)), lambda: (
                torch.from_numpy(np.ascontiguousarray(mp["kernel"], dtype=F)).to(dev),
                torch.from_numpy(np.ascontiguousarray(mp["bias"], dtype=F)).to(dev)))
            f = plane.features.contiguous()
            cells = f.numel() // f.shape[-1]
            out = torch.empty((*f.shape[:-1], self.config.matching_dim), dtype=torch.bfloat16, device=dev)
            ops.match_head(f, plane.valid.contiguous(), cells, f.shape[-1], k, bvec, out,
                           normalize=bool(self.config.normalize_matching_features))
            pred["bev_matching"] = types.FeaturePlane(features=out, valid=plane.valid)
        if self.config.add_confidence:  # bev_mapper.py:292-295
            dev = plane.features.device
            head = params["confidence_head"]["layers_0"]
            cw, cb = self._wcache.lookup(params, ("conf", str(dev)), lambda: (
                torch.from_numpy(np.ascontiguousarray(head["kernel"], dtype=F).reshape(-1)).to(dev),
                float(np.asarray(head["bias"], dtype=F).reshape(-1)[0])))
            f = plane.features.contiguous()
            conf = torch.empty(f.shape[:-1], dtype=torch.float32, device=dev)
            ops.confidence(f, plane.valid.contiguous(), f.numel() // f.shape[-1], f.shape[-1], cw, cb, conf)
            pred["bev_confidence"] = conf
        return pred

    def _apply_per_example(self, variables: Dict, data: Dict, train: bool, debug: bool, is_query: bool) -> Dict:
        """`data['xy_bev']` [B,X,Y,2] with DIFFERENT points per example (`bev_mapper.py:162-166` allows any batch of BEV
        points): the kernels take one point set per launch, so the batch is split into single-example calls and the planes
        are stacked.  (The one caller in the reference, BEVLocalizer, repeats one frustum: `bev_localizer.py:140`.)"""
        import dataclasses
        B = len(data["xy_bev"])

        def take(v, b):
            if isinstance(v, dict):
                return {k: take(x, b) for k, x in v.items()}
            if dataclasses.is_dataclass(v):
                return dataclasses.replace(v, **{f.name: take(getattr(v, f.name), b) for f in dataclasses.fields(v)
                                                 if getattr(v, f.name) is not None})
            if isinstance(v, (np.ndarray, torch.Tensor)) and v.ndim >= 1 and v.shape[0] == B:
                return v[b:b + 1]
            return v
        outs = []
        for b in range(B):
            sub = {k: take(v, b) for k, v in data.items() if k not in ("xyz_grid", "xy_shape", "xy_custom", "staging_slot")}
            sub["xy_bev"] = np.asarray(data["xy_bev"])[b]
            sub["xy_bev_cache"] = False
            p = self.apply(variables, sub, train, debug, is_query)
            outs.append({k: (types.FeaturePlane(p[k].features.clone(), p[k].valid.clone()) if isinstance(p[k], types.FeaturePlane)
                             else p[k].clone()) for k in ("bev_features", "bev_matching", "bev_confidence") if k in p})
        pred: Dict = {}
        for k in outs[0]:
            if isinstance(outs[0][k], types.FeaturePlane):
                pred[k] = types.FeaturePlane(torch.cat([o[k].features for o in outs]), torch.cat([o[k].valid for o in outs]))
            else:
                pred[k] = torch.cat([o[k] for o in outs])
        return pred

    __call__ = apply
