"""B200 implementation behind `snap.models.streetview_encoder.StreetViewEncoder`
(`snap/models/streetview_encoder.py:181-287`): image encoder -> proj MLP -> camera->voxel lift ->
multi-view pooling -> fusion MLP -> FeatureVolume.

v1 pipeline (unfused): crop+ReLU -> proj GEMM (128->160, bias) -> `lift_gather_pool` kernel (projection,
bilinear gather, depth score, softmax pooling; writes the 257-wide statistics rows) -> fusion MLP as two
tcgen05 GEMMs (bias+ReLU / bias+valid-mask epilogues).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _cache, _lib, configs, image_encoder, ops, types

F = np.float32


def fill_lift_params(cfg, V: int, hf: int, wf: int, X: int, Y: int, Z: int, stats_ld: int,
                     xy_paired: bool = False) -> "_lib.LiftParams":
    """Static (shape) part of the lift launch."""
    k_vs = cfg.top_k_view_selection
    if k_vs and V > k_vs:      # view-selection path (`:241-249`)
        if V > _lib.MAX_SELECT_VIEWS or k_vs > _lib.MAX_VIEWS:
            raise NotImplementedError(f"view selection supports top_k <= {_lib.MAX_VIEWS} < V <= {_lib.MAX_SELECT_VIEWS} "
                                      f"(got top_k={k_vs}, V={V})")
    elif V > _lib.MAX_VIEWS:   # all-views path
        raise NotImplementedError(f"all-views lift supports V <= {_lib.MAX_VIEWS} (got V={V} with top_k_view_selection={k_vs})")
    p = _lib.LiftParams()
    p.V, p.Hf, p.Wf = V, hf, wf
    p.D, p.S = cfg.feature_dim, cfg.num_scale_bins
    p.CF = p.D + p.S
    p.X, p.Y, p.Z = X, Y, Z
    dmin, dmax = cfg.depth_min_max
    p.depth_min, p.depth_max = dmin, dmax
    p.inv_log_range = float(F(1.0) / np.log(F(dmax / dmin)).astype(F))
    p.stats_ld = stats_ld
    p.xy_paired = int(xy_paired)
    p.no_variance = int(not cfg.get("fusion_use_variance", True))
    p.add_minmax = int(bool(cfg.get("fusion_add_minmax", False)))
    return p


VIEW_WORDS = C.sizeof(_lib.LiftView) // 4


def pack_views(camera: types.Camera, T_view2scene: types.Transform3D, b: int, stride) -> np.ndarray:
    """Per-scene camera/pose table (`SnapLiftView[V]` as raw 32-bit words), computed on the host in fp32
    in the oracle's operation order: cameras scaled by 1/stride (`streetview_encoder.py:224`) and inverse
    view transforms (`snap/utils/geometry.py:52-56`)."""
    V = camera.f.shape[1]
    cam = camera.scale(np.asarray([1 / stride[1], 1 / stride[0]], dtype=F))
    Tinv = types.Transform3D(R=T_view2scene.R[b], t=T_view2scene.t[b]).inv
    views = (_lib.LiftView * V)()
    for v in range(V):
        lv = views[v]
        for i, x in enumerate(Tinv.R[v].reshape(-1)):
            lv.Rinv[i] = float(x)
        for i in range(3):
            lv.tinv[i] = float(Tinv.t[v, i])
        for i in range(2):
            lv.f[i], lv.c[i], lv.wh[i] = float(cam.f[b, v, i]), float(cam.c[b, v, i]), float(cam.wh[b, v, i])
        if isinstance(camera, types.FisheyeCamera):
            lv.fisheye = 1
            for i in range(3):
                lv.k_radial[i] = float(camera.k_radial[b, v, i])
            lv.tan_half_fov = float(np.tan(F(0.5) * camera.max_fov[b, v]).astype(F))
    return np.frombuffer(bytes(views), dtype=np.int32).copy()


class StreetViewEncoder:
    """Mirror of `snap.models.streetview_encoder.StreetViewEncoder` (`:181-287`)."""

    default_config = staticmethod(configs.streetview_encoder)

    def __init__(self, config=None, dtype=torch.bfloat16):
        self.config = config if config is not None else configs.streetview_encoder()
        c = self.config
        # the per-observation depth_mlp residual (`:263-267`): only exists in the un-weighted branch (`:214-215`); runs on
        # the unfused all-views lift (snapb200_lift_observe -> MLP on the GEMM engine -> snapb200_lift_pool_observations)
        self.has_depth_mlp = c.depth_mlp is not None and not c.do_weighted_fusion
        if self.has_depth_mlp:
            layers_d = tuple(c.depth_mlp.layers or ())
            if not layers_d or layers_d[-1] != c.feature_dim or any(d % 16 for d in layers_d) or \
                    c.depth_mlp.get("activation", "relu") != "relu" or c.depth_mlp.get("apply_input_activation", False):
                raise NotImplementedError("depth_mlp must be a ReLU MLP with layer widths that are multiples of 16 and end in feature_dim")
        # do_weighted_fusion=False (`:262-267` without depth_mlp) runs on the SAME kernels: the sampled maps are the
        # encoder features followed by 32 all-zero scale logits (written by the GEMM engine with the constant operand
        # [I | 0] instead of the proj MLP), which makes the soft-max over the valid views uniform, i.e. [mean | var]
        # are the plain statistics of `pool_multiview_features` (`:153-155`), and the score_max column (= 0) meets a
        # zero row appended to the fusion MLP's first kernel (tests/test_golden.py checks the identity on the oracle).
        self.weighted = bool(c.do_weighted_fusion)
        if not self.weighted and c.image_encoder.output_dim != c.feature_dim:
            raise ValueError("do_weighted_fusion=False samples the encoder features: output_dim must equal feature_dim")
        if tuple(c.fusion.layers) != (256, 128) or c.feature_dim != 128:
            raise NotImplementedError("fusion MLP must be stats->256->128")
        # statistics = [mean | var? | max min? | score_max] (pool_multiview_features :165-177)
        self.stats_dim = c.feature_dim * (1 + int(c.fusion_use_variance) + 2 * int(c.fusion_add_minmax)) + 1
        self.stats_ld = (self.stats_dim + 31) // 32 * 32
        self.default_stats = bool(c.fusion_use_variance) and not c.fusion_add_minmax
        self.dtype = dtype
        self.image_encoder = image_encoder.ImageEncoder(c.image_encoder, dtype)
        self._cache: Dict = {}              # shape-keyed workspaces / staging buffers
        self._wcache = _cache.ParamCache()  # weight banks, per parameter tree

    def clear_cache(self) -> None:
        self._wcache.clear()
        self.image_encoder.clear_cache()

    def _weights(self, params: Dict, device):
        def build():
            bank = image_encoder._WeightBank(device)
            f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1)).to(device)
            fus0_k = np.asarray(params["fusion_mlp"]["Dense_0"]["kernel"], dtype=F)
            if self.weighted:
                proj_k, proj_b = params["proj_mlp"]["Dense_0"]["kernel"], params["proj_mlp"]["Dense_0"]["bias"]
            else:
                d, s = self.config.feature_dim, 160 - self.config.feature_dim
                proj_k, proj_b = np.concatenate([np.eye(d, dtype=F), np.zeros((d, s), F)], 1), np.zeros(d + s, F)
                fus0_k = np.concatenate([fus0_k, np.zeros((1, fus0_k.shape[1]), F)], 0)
            if fus0_k.shape[0] != self.stats_dim:
                raise ValueError(f"fusion_mlp/Dense_0/kernel has {fus0_k.shape[0]} input rows, the configuration needs "
                                 f"{self.stats_dim - (not self.weighted)}")
            depth = []
            if self.has_depth_mlp:
                dp = params["depth_mlp"]
                nlay = len(self.config.depth_mlp.layers)
                if np.asarray(dp["Dense_0"]["kernel"]).shape[0] != self.config.feature_dim + 4:
                    raise ValueError("depth_mlp/Dense_0/kernel must have feature_dim + 4 input rows ([f | log depth | ray])")
                depth = [(bank.add(dp[f"Dense_{i}"]["kernel"], False, k_multiple=32), f32(dp[f"Dense_{i}"]["bias"]))
                         for i in range(nlay)]
            w = dict(bank=bank, depth=depth,
                     proj=bank.add(proj_k, False),
                     fus0=bank.add(fus0_k, False, k_multiple=32),
                     fus1=bank.add(params["fusion_mlp"]["Dense_1"]["kernel"], False),
                     proj_b=f32(proj_b),
                     fus0_b=f32(params["fusion_mlp"]["Dense_0"]["bias"]),
                     fus1_b=f32(params["fusion_mlp"]["Dense_1"]["bias"]),
                     w256=f32(fus0_k[self.stats_dim - 1]))
            bank.finalize()
            return w
        return self._wcache.lookup(params, str(device), build)

    def _buffers(self, device, B, V, H, W, hf, wf, X, Y, Z):
        key = ("buf", str(device), B, V, H, W, hf, wf, X, Y, Z)
        if key not in self._cache:
            N = X * Y * Z
            z = lambda *s, dt=torch.bfloat16: torch.zeros(s, dtype=dt, device=device)
            pin = lambda *s, dt: torch.zeros(s, dtype=dt).pin_memory()
            rows_img = max(V * hf * wf, 128)
            self._cache[key] = dict(
                crop=z(B * rows_img, 128), fimg=z(B, rows_img, 160), stats=z(N, self.stats_ld), hid=z(N, 256),
                volume=None, valid=None,   # [B,N,128] / [B,N]: allocated on first unfused call
                plane=z(B, X * Y, 128), pvalid=z(B, X * Y, dt=torch.uint8), counter=z(B, 16, dt=torch.int32),
                scratch=z(ops.lift_fused_scratch_bytes(), dt=torch.uint8),
                images_host=pin(B, V, H, W, 3, dt=torch.float32), images=z(B, V, H, W, 3, dt=torch.float32),
                xs=None, ys=None)
        return self._cache[key]

    STAGING_SLOTS = 4

    def _staging(self, device, B, Z, slot: int) -> Dict:
        """Per-call inputs (voxel heights, camera / pose tables): pinned host staging + device copies.  Calls are
        asynchronous, so the staging of call i must not be overwritten while its H2D copy is still pending: eager
        calls rotate over STAGING_SLOTS slots and wait on the slot's last copy event before reusing it; a captured
        CUDA graph owns the slot it was captured with (`data['staging_slot']`) and re-reads it at every replay."""
        key = ("stg", str(device), B, Z, slot)
        if key not in self._cache:
            pin = lambda *s, dt: torch.zeros(s, dtype=dt).pin_memory()
            z = lambda *s, dt: torch.zeros(s, dtype=dt, device=device)
            self._cache[key] = dict(zs_host=pin(B, Z, dt=torch.float32), zs=z(B, Z, dt=torch.float32),
                                    views_host=pin(B, _lib.MAX_SELECT_VIEWS * VIEW_WORDS, dt=torch.int32),
                                    views=z(B, _lib.MAX_SELECT_VIEWS * VIEW_WORDS, dt=torch.int32),
                                    centers_host=pin(B, _lib.MAX_SELECT_VIEWS * 3, dt=torch.float32),
                                    centers=z(B, _lib.MAX_SELECT_VIEWS * 3, dt=torch.float32), event=None)
        return self._cache[key]

    def stage_inputs(self, data: Dict, buf: Dict, stride) -> None:
        """Host-only: write this batch's voxel heights and camera/pose tables into the pinned staging buffers
        (a captured CUDA graph re-reads them at every replay)."""
        xs, ys, zs = data["xyz_grid"]
        B = zs.shape[0]
        buf["zs_host"].copy_(torch.from_numpy(np.ascontiguousarray(zs, dtype=F)))
        for b in range(B):
            pack = pack_views(data["camera"], data["T_view2scene"], b, stride)
            buf["views_host"][b, : len(pack)].copy_(torch.from_numpy(pack))
            cen = np.ascontiguousarray(data["T_view2scene"].t[b], dtype=F).reshape(-1)   # camera centres (`:131`)
            buf["centers_host"][b, : len(cen)].copy_(torch.from_numpy(cen))

    def upload_staging(self, variables: Dict, data: Dict, device=None) -> None:
        """Input-pipeline hook: stage this batch's voxel heights and camera / pose tables and enqueue their H2D copies
        on the CURRENT stream (e.g. the copy stream that also uploads the images), into staging slot
        `data['staging_slot']`.  A later `apply` with `data['staging_uploaded'] = True` (same slot) then contains no
        host-to-device copy at all, so an upload running beside it never delays its first kernels."""
        params = variables["params"] if "params" in variables else variables
        images = data["images"]
        B, V, H, W, _ = images.shape
        dev = torch.device(device) if device is not None else (images.device if images.is_cuda else torch.device("cuda"))
        zs = data["xyz_grid"][2]
        enc_plan = self.image_encoder.plan(params["image_encoder"], B * V, H, W, dev)
        stg = self._staging(dev, B, zs.shape[1], data["staging_slot"])
        self.stage_inputs(data, stg, enc_plan.strides[-1])
        stg["zs"].copy_(stg["zs_host"], non_blocking=True)
        stg["views"].copy_(stg["views_host"], non_blocking=True)
        stg["centers"].copy_(stg["centers_host"], non_blocking=True)

    def uses_view_selection(self, V: int) -> bool:
        """`:241`: the top-k view-selection path runs iff the scene has more views than top_k_view_selection."""
        k_vs = self.config.top_k_view_selection
        return bool(k_vs) and V > k_vs

    def apply(self, variables: Dict, data: Dict, train: bool = False, debug: bool = False,
              fused: bool = False) -> Dict:
        """`fused=True` runs the whole lift as one kernel and returns 'feature_plane' (bev_mapper.py:56-88 applied)
        instead of materialising 'feature_volume' (which then is absent from the result); the fused kernel only
        covers the all-views path (V <= top_k_view_selection, V <= 4)."""
        if train:
            raise NotImplementedError("training (backward kernels) is a 'next' row of SURVEY.md §8(f)")
        params = variables["params"] if "params" in variables else variables
        cfg = self.config
        images = data["images"]
        B, V, H, W, _ = images.shape
        select = self.uses_view_selection(V)
        if fused and select:
            raise ValueError("the fused lift kernel has no view selection: call with fused=False when V > top_k_view_selection")
        dev = images.device if isinstance(images, torch.Tensor) and images.is_cuda else torch.device("cuda")
        xs, ys, zs = data["xyz_grid"]          # xs [X], ys [Y] (NumPy fp32), zs [B, Z]
        X, Y, Z = len(xs), len(ys), zs.shape[1]
        paired = "xy_shape" in data            # data['xy_bev'] that is not a separable grid: xs, ys per column [X*Y]
        if paired:
            X, Y = data["xy_shape"]
            assert len(xs) == X * Y and len(ys) == X * Y
        enc_plan = self.image_encoder.plan(params["image_encoder"], B * V, H, W, dev)
        hf, wf = enc_plan.cropped_shapes()[-1]
        stride = enc_plan.strides[-1]
        buf = self._buffers(dev, B, V, H, W, hf, wf, X, Y, Z)
        custom = bool(paired or data.get("xy_custom"))   # data['xy_bev'] / data['xyz_query'] instead of the mapper's grid
        if buf["xs"] is None or (custom and buf.get("xs_src") is not xs) or (not custom and buf.get("xs_custom")):
            buf["xs"], buf["ys"] = torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev)
            buf["xs_src"], buf["xs_custom"] = xs, custom
        capturing = torch.cuda.is_current_stream_capturing()
        slot = data.get("staging_slot")
        if slot is None:
            slot = 0
            if not capturing:
                slot = self._slot_counter = (getattr(self, "_slot_counter", -1) + 1) % self.STAGING_SLOTS
        stg = self._staging(dev, B, Z, slot)
        pre_uploaded = bool(data.get("staging_uploaded", False))   # see `upload_staging`
        if not pre_uploaded:
            if not capturing and stg["event"] is not None:
                stg["event"].synchronize()      # the slot's previous H2D copy has drained
            self.stage_inputs(data, stg, stride)
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images, dtype=F))
        if not images.is_cuda:   # host images: pinned -> direct async H2D, pageable -> via the pinned staging buffer
            staged = not images.is_pinned()
            if staged:
                if buf.get("images_event") is not None:
                    buf["images_event"].synchronize()   # previous upload from the staging buffer has drained
                buf["images_host"].copy_(images)
                images = buf["images_host"]
            buf["images"].copy_(images, non_blocking=True)
            if staged and not capturing:
                buf["images_event"] = torch.cuda.Event()
                buf["images_event"].record()
            images = buf["images"]
        if not pre_uploaded:
            stg["zs"].copy_(stg["zs_host"], non_blocking=True)
            stg["views"].copy_(stg["views_host"], non_blocking=True)
            if select:
                stg["centers"].copy_(stg["centers_host"], non_blocking=True)
            if not capturing:
                stg["event"] = torch.cuda.Event()
                stg["event"].record()

        pyr = data.get("image_feature_pyr")
        if pyr is None:
            pyr = self.image_encoder.apply({"params": params["image_encoder"]}, images.reshape(B * V, H, W, 3), train)
        full = pyr.uncropped[-1]               # [B*V, Hs, Ws, 128] un-cropped finest FPN level
        Hs, Ws = full.shape[1], full.shape[2]
        wts = self._weights(params, dev)
        bank = wts["bank"]
        bank.run()
        Bm = bank.b_mats
        N = X * Y * Z
        lp = fill_lift_params(cfg, V, hf, wf, X, Y, Z, self.stats_ld, paired)
        if (fused or select) and not self.default_stats:
            raise NotImplementedError("fusion_add_minmax / fusion_use_variance=False run on the unfused all-views lift only")
        if self.has_depth_mlp and (fused or select or not self.default_stats):
            raise NotImplementedError("depth_mlp runs on the unfused all-views lift with the default statistics "
                                      "(V <= top_k_view_selection, fused=False)")
        dbg = {}
        if not fused and buf["volume"] is None:
            buf["volume"] = torch.zeros((B, N, 128), dtype=torch.bfloat16, device=dev)
            buf["valid"] = torch.zeros((B, N), dtype=torch.uint8, device=dev)
        # The batched path: ONE crop + ONE proj GEMM + ONE launch of the warp-specialised lift for all scenes
        # (`SNAPB200_LIFT_V=1` keeps the first-generation per-scene kernel for A/B measurements).
        batched = fused and os.environ.get("SNAPB200_LIFT_V", "2") != "1"
        chunk = max(1, min(32 // V, (1 << 18) // (X * Y)))   # scenes per launch: B * V <= 32 views, B * X * Y <= 2^18 columns
        rows_img = buf["fimg"].shape[1]
        if batched:
            if rows_img == V * hf * wf:
                ops.crop_relu(full, B * V, Hs, Ws, 128, hf, wf, self.weighted, buf["crop"])
                ops.gemm(buf["crop"], Bm[wts["proj"]], buf["fimg"].view(B * rows_img, 160), m_rows=B * rows_img,
                         bias=wts["proj_b"])
            else:   # tiny feature maps (a scene's rows are padded to the GEMM's 128-row tile): per scene
                for b in range(B):
                    ops.crop_relu(full[b * V:(b + 1) * V], V, Hs, Ws, 128, hf, wf, self.weighted, buf["crop"])
                    ops.gemm(buf["crop"], Bm[wts["proj"]], buf["fimg"][b], m_rows=V * hf * wf, bias=wts["proj_b"])
            for ci, b0 in enumerate(range(0, B, chunk)):
                b1 = min(B, b0 + chunk)
                ops.lift_fused_batched(lp, b1 - b0, stg["views"][b0:b1], buf["fimg"][b0:b1], buf["xs"], buf["ys"], stg["zs"][b0:b1],
                                       Bm[wts["fus0"]], wts["w256"], wts["fus0_b"], Bm[wts["fus1"]], wts["fus1_b"],
                                       buf["plane"][b0:b1], buf["pvalid"][b0:b1], buf["counter"][ci], buf["scratch"])
            buf["lift_launches"] = -(-B // chunk)
        for b in range(0 if not batched else B, B):
            # proj_mlp: ReLU -> Dense(128 -> 160) on the cropped finest level (`:228-230`); un-weighted fusion: the crop
            # itself, widened by zero logits (see __init__)
            ops.crop_relu(full[b * V:(b + 1) * V], V, Hs, Ws, 128, hf, wf, self.weighted, buf["crop"])
            ops.gemm(buf["crop"], Bm[wts["proj"]], buf["fimg"][b], m_rows=V * hf * wf, bias=wts["proj_b"])
            if fused:
                ops.lift_fused(lp, stg["views"][b], buf["fimg"][b], buf["xs"], buf["ys"], stg["zs"][b],
                               Bm[wts["fus0"]], wts["w256"], wts["fus0_b"], Bm[wts["fus1"]], wts["fus1_b"],
                               buf["plane"][b], buf["pvalid"][b], buf["counter"][b], buf["scratch"])
                continue
            dv = dt = di = None
            Kd = cfg.top_k_view_selection if select else V
            if debug:
                dv = torch.zeros((N, Kd), dtype=torch.uint8, device=dev)
                dt = torch.zeros((N, Kd, 2), dtype=torch.int32, device=dev)
                dbg.setdefault("vis", []).append(dv)
                dbg.setdefault("taps", []).append(dt)
                if select:
                    di = torch.zeros((N, Kd), dtype=torch.int32, device=dev)
                    dbg.setdefault("view_indices", []).append(di)
            if self.has_depth_mlp:   # un-weighted branch with the per-observation residual (`:263-267`)
                ob = self._cache.get(("obs", str(dev), N, V))
                if ob is None:
                    widths = [160] + list(cfg.depth_mlp.layers)
                    ob = self._cache[("obs", str(dev), N, V)] = dict(
                        x=[torch.zeros((image_encoder._round_up(N * V, 128), wd), dtype=torch.bfloat16, device=dev) for wd in widths],
                        vis=torch.zeros((N, V), dtype=torch.uint8, device=dev))
                ops.lift_observe(lp, stg["views"][b], buf["fimg"][b], buf["xs"], buf["ys"], stg["zs"][b], ob["x"][0], ob["vis"])
                for i, (wi, bi) in enumerate(wts["depth"]):      # layers.MLP: Dense -> (relu -> Dense)*
                    ops.gemm(ob["x"][i], Bm[wi], ob["x"][i + 1], m_rows=N * V, seg_k=ob["x"][i].shape[1], bias=bi,
                             relu=i + 1 < len(wts["depth"]))
                ops.lift_pool_observations(V, N, ob["x"][0], ob["x"][-1], ob["vis"], buf["stats"], buf["valid"][b])
                if debug:
                    dv.copy_(ob["vis"])
            elif select:   # V > top_k: view selection + selective sampling (`:241-249`)
                ops.lift_select_pool(lp, cfg.top_k_view_selection, cfg.get("max_view_distance"), stg["views"][b],
                                     stg["centers"][b], buf["fimg"][b], buf["xs"], buf["ys"], stg["zs"][b],
                                     buf["stats"], buf["valid"][b], di, dv, dt)
            else:
                ops.lift_gather_pool(lp, stg["views"][b], buf["fimg"][b], buf["xs"], buf["ys"], stg["zs"][b],
                                     buf["stats"], buf["valid"][b], dv, dt)
            # fusion MLP 257 -> 256 -> 128 (`:281`), zero where invalid (`:282`)
            ops.gemm(buf["stats"], Bm[wts["fus0"]], buf["hid"], m_rows=N, seg_k=self.stats_ld, bias=wts["fus0_b"], relu=True)
            ops.gemm(buf["hid"], Bm[wts["fus1"]], buf["volume"][b], m_rows=N, bias=wts["fus1_b"],
                     row_mask=buf["valid"][b])
        pred = {"image_feature_pyramid": pyr}
        # what the lift's backward needs of this call (references to the live buffers: valid until the next call with the
        # same shapes): `snap_b200.localizer_trainer` turns it into one `SceneContext` per example
        pred["lift_context"] = dict(lp=lp, views=stg["views"], centers=stg["centers"], fimg=buf["fimg"], crop=buf["crop"],
                                    xs=buf["xs"], ys=buf["ys"], zs=stg["zs"], rows_img=rows_img, V=V, hf=hf, wf=wf,
                                    full=full, Hs=Hs, Ws=Ws, relu_crop=self.weighted,
                                    crop_per_scene=not (batched and rows_img == V * hf * wf) and B > 1)
        if self.weighted:                      # `:229-230`
            pred["scores_images"] = buf["fimg"][:, :V * hf * wf].view(B, V, hf, wf, 160)[..., 128:]
        if fused:
            pred["feature_plane"] = types.FeaturePlane(features=buf["plane"].view(B, X, Y, 128),
                                                       valid=buf["pvalid"].view(B, X, Y))
        else:
            pred["feature_volume"] = types.FeatureVolume(features=buf["volume"].view(B, X, Y, Z, 128),
                                                         valid=buf["valid"].view(B, X, Y, Z))
        if debug:
            pred["debug"] = {k: torch.stack(v) for k, v in dbg.items()}
            pred["debug"]["f_proj_images"] = buf["fimg"][:, :V * hf * wf].view(B, V, hf, wf, 160)
        return pred

    __call__ = apply
