"""Backward of the sampling localizer's loss (`snap/models/bev_localizer.py:156-160,183-216,244-262`) from the pose scores
down to the two matching planes: what `jax.grad` of `mean_b(nll_b)` computes w.r.t. `plane_q.features`,
`plane_map.features` and `temperature` (the sampled poses are integer draws and carry no gradient).

    scores [B,P1]  --loc_nll_backward-->            dscores, dtemperature
    dscores        --loc_pose_scoring_backward-->   dsim bf16 [B,N,H*W] (ReLU-masked cotangent of the similarities; each
                                                    point's map is accumulated in shared memory and written once)
    d f_q [B,N,D]  = dsim f_m      the tcgen05 GEMM engine (B operand = f_m^T via `wt_segments`)
    d f_m [B,HW,D] = dsim^T f_q    the split-K weight-gradient kernel over <= 1024-column slices of dsim
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import image_encoder, ops


class LocalizerLossBackward:
    def __init__(self, device):
        self.dev = device
        self._buf: Dict = {}

    def _buffers(self, B: int, N: int, HW: int, D: int, P1: int) -> Dict:
        key = (B, N, HW, D, P1)
        if key not in self._buf:
            Np = image_encoder._round_up(max(N, 128), 128)
            z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=self.dev)
            self._buf[key] = dict(Np=Np, dscores=z(B, P1), dtemp=z(B), dsim=z(B, Np, HW, dt=torch.bfloat16),
                                  fq=z(B, Np, D, dt=torch.bfloat16), fmT=z(D, HW, dt=torch.bfloat16),
                                  dfq=z(B, Np, D, dt=torch.bfloat16), dfm=z(B, HW, D))
        return self._buf[key]

    def backward(self, maps, f_p_q: torch.Tensor, map_features: torch.Tensor, q_xy_p: torch.Tensor,
                 valid_j: Optional[torch.Tensor], poses: torch.Tensor, scores: torch.Tensor, cell_size: float,
                 mask_out_of_bounds: bool, clip_negative_scores: bool = True,
                 remove: Optional[Sequence[float]] = None, dr_samples: Optional[torch.Tensor] = None,
                 dt_samples: Optional[torch.Tensor] = None, example_weights: Optional[torch.Tensor] = None):
        """maps = the forward's `pose_estimation.SimilarityMaps`; f_p_q bf16 [B,N,D]; map_features bf16 [B,H,W,D]; poses f32
        [B,P1,3] (ground truth first) and scores f32 [B,P1] as scored by the forward; dr / dt = the per-sample errors of
        `loc_nll` when `remove` (threshold_remove_accurate_poses) is set.
        Returns (d f_q bf16 [B,N,D], d f_m f32 [B,H*W,D], dtemperature f32 [B])."""
        if getattr(maps, "from_confidence", False):
            # `bev_localizer.py:165-169`: sim_points is multiplied by masked_softmax(bev_confidence), whose gradient flows
            # into the trainable confidence head; this plan treats the point weights as constants
            raise NotImplementedError("add_confidence_query: the gradient of the query confidences is not implemented")
        B, N, D = f_p_q.shape
        H, W = maps.H, maps.W
        HW, P1 = H * W, scores.shape[1]
        if D % 16 or HW % 32:
            raise NotImplementedError("matching_dim must be a multiple of 16 and H * W of 32")
        buf = self._buffers(B, N, HW, D, P1)
        Np = buf["Np"]
        ops.loc_nll_backward(scores.contiguous(), remove, dr_samples, dt_samples, buf["dscores"], buf["dtemp"])
        if example_weights is not None:
            # the kernel differentiates mean_b(nll_b); `trainer.py:221` means over batch['batch_mask'] instead: example b
            # gets the weight mask_b * B / sum(mask) (f32 [B], device)
            buf["dscores"].mul_(example_weights[:, None])
            buf["dtemp"].mul_(example_weights)
        ops.loc_pose_scoring_backward(maps.sim, maps.point_scale, q_xy_p.contiguous(), valid_j, poses.contiguous(),
                                      buf["dscores"], H, W, cell_size, mask_out_of_bounds, clip_negative_scores, buf["dsim"])
        buf["fq"][:, :N].copy_(f_p_q)                      # rows N..Np stay zero (split-K kernel: M multiple of 16)
        fm = map_features.contiguous().view(B, HW, D)
        for b in range(B):
            ops.wt_segments(fm[b], HW, D, 1, buf["fmT"])                                    # f_m^T [D, HW]
            ops.gemm(buf["dsim"][b], buf["fmT"], buf["dfq"][b], m_rows=Np, seg_k=HW)         # d f_q = dsim f_m
            for c0 in range(0, HW, 1024):                                                   # d f_m = dsim^T f_q
                kc = min(1024, HW - c0)
                ops.dense_wgrad(buf["dsim"][b][:, c0:c0 + kc], buf["fq"][b], Np, kc, D, buf["dfm"][b][c0:c0 + kc], None)
        return buf["dfq"][:, :N], buf["dfm"], buf["dtemp"]
