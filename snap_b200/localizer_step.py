"""Backward of one localisation training step with FROZEN image encoders (`freeze_params_reg_exp` on the encoders; the
trainable arrays are `proj_mlp`, `fusion_mlp`, `matching_proj` of the map / query BEV mappers and the temperature):
the per-module launch plans chained in the order `jax.grad` of `snap/models/bev_localizer.py` traverses them.

    NLL (`bev_localizer.py:244-262`)  -> `LocalizerLossBackward`            d f_q, d f_m, d temperature
    per example, map side:               `MatchingHeadBackward`              d matching_proj, cotangent of `bev_features`
                                         [`fusion_backward` if an aerial plane was fused in (its encoder is frozen)]
                                         `LiftBackward.scene_backward`       d fusion_mlp, d proj_mlp
    per example, query side:             the same on the field-of-view points (`data['xy_bev']`, one BEV column per point)

What the forward of one scene has to leave behind is listed in `SceneContext`; nothing per voxel is kept (the lift backward
recomputes the scene from its projected feature maps).  The step driver proper -- extracting the contexts from
`BEVLocalizer.apply`, the gradient mean over ranks and Adam on shared masters -- is GPU-side plumbing on top of this."""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional, Sequence

import torch

from . import image_encoder, localizer_train, streetview_train


@dataclasses.dataclass
class SceneContext:
    """What one scene's BEV forward leaves for the backward (device tensors; `lp` = the lift's parameter struct)."""
    lp: Any
    views: Optional[torch.Tensor]            # packed camera / pose table of the scene
    fimg: torch.Tensor                       # bf16 [V*hf*wf, 160] projected feature maps (proj MLP output)
    crop: torch.Tensor                       # bf16 [V*hf*wf, 128] relu(cropped finest FPN level)
    xs: Optional[torch.Tensor]
    ys: Optional[torch.Tensor]
    zs: Optional[torch.Tensor]
    plane: torch.Tensor                      # bf16 [cells, 128] street-view feature plane
    plane_valid: torch.Tensor                # u8 [cells]
    aerial_plane: Optional[torch.Tensor] = None     # bf16 [cells, 128]: modality fusion 'max' with an (all-valid) aerial plane
    fused_plane: Optional[torch.Tensor] = None      # bf16 [cells, 128] = `bev_features` when an aerial plane was fused in
    top_k: Optional[int] = None              # view-selection path (V > top_k)
    view_centers: Optional[torch.Tensor] = None
    max_view_distance: Optional[float] = None


class FrozenEncoderBackward:
    def __init__(self, mapper_params: Dict, device, query_mapper_params: Optional[Dict] = None):
        """mapper_params: the Flax tree of the map `bev_mapper` ('streetview_encoder' {proj_mlp, fusion_mlp}, 'matching_proj');
        query_mapper_params: the tree of `bev_mapper_query` if the model has a separate query mapper
        (`bev_localizer.py:93-100`), else the query side shares -- and adds its gradients to -- the map mapper's arrays."""
        self.dev = device
        self.loc = localizer_train.LocalizerLossBackward(device)
        mk = lambda p: (streetview_train.LiftBackward(p["streetview_encoder"], device),
                        streetview_train.MatchingHeadBackward(p["matching_proj"], device))
        self.lift_map, self.head_map = mk(mapper_params)
        self.shared = query_mapper_params is None
        self.lift_q, self.head_q = (self.lift_map, self.head_map) if self.shared else mk(query_mapper_params)

    def _side(self, lift, head, ctx: SceneContext, dmatch: torch.Tensor) -> torch.Tensor:
        """One scene of one side: dmatch bf16 [>= cells, 32] -> parameter gradients (accumulated), returns dcrop."""
        cells = ctx.plane_valid.numel()
        Cp = image_encoder._round_up(cells, 16)                  # split-K kernels need a multiple of 16 rows
        fused = ctx.aerial_plane is not None
        head_plane = ctx.fused_plane if fused else ctx.plane
        head_valid = torch.ones(cells, dtype=torch.uint8, device=self.dev) if fused else ctx.plane_valid   # bev_mapper.py:208-211
        if Cp != cells:                                           # zero rows: they carry neither features nor cotangents
            pad = lambda t, c, dt: torch.cat([t.reshape(cells, c), torch.zeros((Cp - cells, c), dtype=dt, device=self.dev)])
            head_plane = pad(head_plane, head_plane.shape[-1], torch.bfloat16).contiguous()
            head_valid = pad(head_valid, 1, torch.uint8).reshape(-1).contiguous()
            dmatch = pad(dmatch[:cells], 32, torch.bfloat16).contiguous()
        dplane = head.backward(head_plane, head_valid, dmatch, accumulate=True)[:cells]
        if fused:
            dplane, _ = head.fusion_backward(ctx.plane, ctx.plane_valid, ctx.aerial_plane, dplane.contiguous())
        return lift.scene_backward(ctx.lp, ctx.views, ctx.fimg, ctx.crop, ctx.xs, ctx.ys, ctx.zs, None, None,
                                   dplane.contiguous(), top_k=ctx.top_k, view_centers=ctx.view_centers,
                                   max_view_distance=ctx.max_view_distance)

    def backward(self, maps, f_p_q: torch.Tensor, map_features: torch.Tensor, q_xy_p: torch.Tensor,
                 valid_j: Optional[torch.Tensor], poses: torch.Tensor, scores: torch.Tensor, cell_size: float,
                 mask_out_of_bounds: bool, clip_negative_scores: bool, remove: Optional[Sequence[float]],
                 dr_samples: Optional[torch.Tensor], dt_samples: Optional[torch.Tensor],
                 map_scenes: List[SceneContext], query_scenes: List[SceneContext],
                 example_weights: Optional[torch.Tensor] = None) -> Dict:
        """Arguments up to `dt_samples`: those of `LocalizerLossBackward.backward`; then one `SceneContext` per example
        for the map and the query side.  Returns {'bev_mapper': grads tree, ['bev_mapper_query': grads tree,]
        'temperature': f32 [], 'encoder_cotangents': {'map': [dcrop per example], 'query': [...]}} (device tensors for the
        scalars, host arrays in the trees)."""
        B = f_p_q.shape[0]
        for m in {id(self.lift_map): self.lift_map, id(self.lift_q): self.lift_q}.values():
            m.zero_grads()
        for m in {id(self.head_map): self.head_map, id(self.head_q): self.head_q}.values():
            m.zero_grads()
        dfq, dfm, dtemp = self.loc.backward(maps, f_p_q, map_features, q_xy_p, valid_j, poses, scores, cell_size,
                                            mask_out_of_bounds, clip_negative_scores, remove, dr_samples, dt_samples,
                                            example_weights=example_weights)
        enc = {"map": [], "query": []}
        for b in range(B):
            enc["map"].append(self._side(self.lift_map, self.head_map, map_scenes[b], dfm[b].to(torch.bfloat16)).clone())
            enc["query"].append(self._side(self.lift_q, self.head_q, query_scenes[b], dfq[b].contiguous()).clone())
        tree = lambda lift, head: {"streetview_encoder": lift.grads_tree(),
                                   "matching_proj": {k: v.cpu().numpy().copy() for k, v in head.g.items()}}
        out = {"bev_mapper": tree(self.lift_map, self.head_map), "temperature": dtemp.sum(), "encoder_cotangents": enc}
        if not self.shared:
            out["bev_mapper_query"] = tree(self.lift_q, self.head_q)
        return out
