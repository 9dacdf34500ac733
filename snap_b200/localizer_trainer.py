"""One data-parallel training step of the localisation model (`snap/trainer.py:165-295` around
`snap/models/bev_localizer.py`) with FROZEN image encoders: forward (`BEVLocalizer.apply` with the ground-truth pose
prepended to the samples) -> NLL loss (`bev_localizer.py:244-262`) -> backward (`localizer_step.FrozenEncoderBackward`)
-> gradient mean over ranks (`jax.lax.pmean`, `trainer.py:231-234`) -> optional global-norm clipping (`:238-239`)
-> Adam (`:242-245`, optax.adam on fp32 masters) -> skip of the whole update when a gradient is not finite (`:260-276`).

Trainable arrays: `proj_mlp`, `fusion_mlp` and `matching_proj` of the map / query BEV mappers and `temperature`; the
image encoders are frozen (their cotangent is returned per scene for `encoder_train.TrunkTrainer.backward`).  The masters
live in ONE flat fp32 device buffer laid out like the gradient bucket (`parallel.GradBucket`), so the collective, the
finite check, the clipping and Adam are one pass each over ~0.5 MB; after an applied step the Flax parameter tree is
rebuilt from the masters with NEW dict objects for the trainable sub-trees only -- the frozen `image_encoder` /
`aerial_encoder` sub-trees keep their identity, so the forward's per-tree weight caches (`_cache.ParamCache`) re-upload
just the small heads."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import bev_localizer, encoder_train, localizer_step, ops, parallel, types

F = np.float32


class EncoderTraining:
    """The image encoder of the (shared) BEV mapper in training mode: `encoder_train.TrunkTrainer` over ALL images of a step
    (map views and query views in one batch, so there is one set of masters and one gradient per array), padded like
    `pad_to_multiple` (`image_encoder.py:32-39`: a full stride when already aligned).  Parameters and gradients stay on
    the device: `leaves` lists (path, parameter tensor, gradient tensor) for the optimiser; the parameter tensors are the
    ones the training forward reads (fp32 conv masters that every forward re-standardises, GroupNorm scale / bias)."""

    def __init__(self, enc_params: Dict, n_img: int, H: int, W: int, device):
        s = 32
        self.n, self.H, self.W = n_img, H, W
        self.Hp, self.Wp = H + (s - H % s), W + (s - W % s)
        self.trunk = t = encoder_train.TrunkTrainer(enc_params, n_img, self.Hp, self.Wp, device)
        self.padded = torch.zeros((n_img, self.Hp, self.Wp, 3), dtype=torch.float32, device=device)
        self.Hs, self.Ws = self.Hp // 4, self.Wp // 4
        self.dfin = torch.zeros((n_img, self.Hs, self.Ws, 128), dtype=torch.bfloat16, device=device)
        leaves = [(("encoder", "root_block", "conv_root", "kernel"), t.root_master, t.g_root)]
        for u in t.units:
            ut = u["t"]
            for name in ut.master:
                leaves.append((("encoder", *u["path"], name, "kernel"), ut.master[name], ut.g[name]))
            for g in ("gn1", "gn2", "gn3"):
                leaves.append((("encoder", *u["path"], g, "scale"), ut.gn[g][0], ut.ggn[g][0]))
                leaves.append((("encoder", *u["path"], g, "bias"), ut.gn[g][1], ut.ggn[g][1]))
        off = 0
        for level, (L, (_, k, cout, _, _)) in enumerate(zip(t.fpn.lv, t.fpn.bank.entries)):
            leaves.append((("decoder", f"{level}_skip_conv", "kernel"), t.fpn.bank.master[off: off + k * cout].view(k, cout),
                           L["g"]["kernel"]))
            off += k * cout
            leaves.append((("decoder", f"{level}_skip_norm", "scale"), L["scale"], L["g"]["scale"]))
            leaves.append((("decoder", f"{level}_skip_norm", "bias"), L["bias"], L["g"]["bias"]))
        self.leaves = leaves

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """images f32 [n, H, W, 3] (device) -> un-cropped finest FPN level bf16 [n, Hp/4, Wp/4, 128]."""
        self.padded[:, : self.H, : self.W].copy_(images)
        fin = self.trunk.forward(self.padded)
        return fin[: self.n * self.Hs * self.Ws].view(self.n, self.Hs, self.Ws, 128)

    def pyramid(self, feats: torch.Tensor, hf: int, wf: int) -> types.FeatureImagePyramid:
        stride = np.asarray([self.Hp / self.Hs, self.Wp / self.Ws])
        return types.FeatureImagePyramid(features=[feats[:, :hf, :wf]], strides=[stride], uncropped=[feats])

    def backward(self, dcrops: List[torch.Tensor], V_of: List[int], hf: int, wf: int) -> None:
        """dcrops[i] bf16 [>= V_i*hf*wf, 128]: cotangent of the cropped finest level of the i-th group of V_i consecutive
        images (one scene of one side); the padding rows / columns get zero."""
        self.dfin.zero_()
        i0 = 0
        for d, V in zip(dcrops, V_of):
            self.dfin[i0:i0 + V, :hf, :wf].copy_(d[: V * hf * wf].view(V, hf, wf, 128))
            i0 += V
        self.trunk.backward(self.dfin.view(-1, 128))


def clip_scale(norm: float, max_norm: float) -> float:
    """`jax.example_libraries.optimizers.clip_grads` (the reference's `trainer.py:30,239`): gradients are multiplied by
    max_norm / norm when the global l2 norm is not below max_norm."""
    return 1.0 if norm < max_norm else max_norm / norm


class LocalizerTrainer:
    LEAVES = (("streetview_encoder", "proj_mlp", "Dense_0", "kernel"), ("streetview_encoder", "proj_mlp", "Dense_0", "bias"),
              ("streetview_encoder", "fusion_mlp", "Dense_0", "kernel"), ("streetview_encoder", "fusion_mlp", "Dense_0", "bias"),
              ("streetview_encoder", "fusion_mlp", "Dense_1", "kernel"), ("streetview_encoder", "fusion_mlp", "Dense_1", "bias"),
              ("matching_proj", "kernel"), ("matching_proj", "bias"))

    def __init__(self, model: bev_localizer.BEVLocalizer, params: Dict, lr: Union[float, Callable[[int], float]] = 5e-5,
                 max_grad_norm: Optional[float] = None, device=None, group=None, train_encoder: bool = False):
        """params: the Flax tree {'bev_mapper', ['bev_mapper_query',] 'temperature'} (host arrays); lr: a constant or the
        schedule `step -> learning rate` (`configs/train_localization.py:86-92`).
        train_encoder: also train the street-view image encoder (the reference's default for this model: no
        `freeze_params_reg_exp` in `configs/train_localization.py`): its training forward / backward is
        `encoder_train.TrunkTrainer` over the map and query images of the step; the ~23.5 M encoder gradients join the flat
        bucket of the collective.  The aerial encoder, if any, stays frozen."""
        c = model.config
        if c.add_confidence_query:
            # the confidence head multiplies the point similarities (`bev_localizer.py:165-169`) and would need its own
            # gradient: refuse instead of training with an incomplete gradient
            raise NotImplementedError("add_confidence_query: the confidence head's backward is not implemented")
        if not c.add_temperature:
            raise NotImplementedError("the step driver trains the temperature (`add_temperature=True`, the default)")
        self.model, self.dev = model, torch.device(device if device is not None else "cuda")
        self.lr, self.max_grad_norm, self.step, self.skipped_steps = lr, max_grad_norm, 0, 0
        self.sides = ["bev_mapper"] + (["bev_mapper_query"] if model.bev_mapper_query is not None else [])
        self.params = params
        self.paths: List[Tuple[str, ...]] = [(s, *leaf) for s in self.sides for leaf in self.LEAVES] + [("temperature",)]
        get = lambda path: np.asarray(self._get(params, path), dtype=F)
        self.masters = parallel.GradBucket([get(p).shape or (1,) for p in self.paths], self.dev, group)
        self.bucket = parallel.GradBucket(self.masters.shapes, self.dev, group)
        for v, path in zip(self.masters.views, self.paths):
            v.copy_(torch.from_numpy(get(path).reshape(v.shape).copy()))
        self.mom = (torch.zeros_like(self.masters.flat), torch.zeros_like(self.masters.flat))
        self.bwd = localizer_step.FrozenEncoderBackward(
            params["bev_mapper"], self.dev, params.get("bev_mapper_query") if len(self.sides) == 2 else None)
        self._crop: Dict = {}
        self.train_encoder, self.enc, self.enc_bucket, self.enc_mom = train_encoder, None, None, None
        if train_encoder and len(self.sides) == 2:
            raise NotImplementedError("train_encoder with a separate query mapper (two encoders) is not wired")

    def _encoder(self, n_img: int, H: int, W: int) -> EncoderTraining:
        if self.enc is None or (self.enc.n, self.enc.H, self.enc.W) != (n_img, H, W):
            if self.enc is not None:
                raise ValueError("the encoder's training buffers are built for one input shape per trainer")
            enc_params = self.params["bev_mapper"]["streetview_encoder"]["image_encoder"]
            self.enc = EncoderTraining(enc_params, n_img, H, W, self.dev)
            self.enc_bucket = parallel.GradBucket([tuple(p.shape) for _, p, _ in self.enc.leaves], self.dev, self.masters.group)
            self.enc_mom = [(torch.zeros_like(p), torch.zeros_like(p)) for _, p, _ in self.enc.leaves]
        return self.enc

    @staticmethod
    def _get(tree: Dict, path):
        for k in path:
            tree = tree[k]
        return tree

    # ---------------------------------------------------------------------------------------------------------------
    def _contexts(self, pred_side: Dict, key: str) -> List[localizer_step.SceneContext]:
        """One `SceneContext` per example from what `BEVMapper.apply` left behind (references, no copies)."""
        sv = pred_side["streetview"]
        lc, plane = sv["lift_context"], sv["feature_plane"]
        B = plane.features.shape[0]
        cells = plane.valid[0].numel()
        rows, lp = lc["rows_img"], lc["lp"]
        aerial = pred_side.get("aerial", {}).get("feature_plane")
        fusedp = pred_side["bev_features"] if aerial is not None else None
        sel = self.model.bev_mapper.streetview_encoder if key == "bev_mapper" else \
            (self.model.bev_mapper_query or self.model.bev_mapper).streetview_encoder
        select = sel.uses_view_selection(lc["V"])
        out = []
        for b in range(B):
            if lc["crop_per_scene"]:     # the forward reused one crop buffer per scene: cut this scene's crop again
                ck = (key, b, rows)
                if ck not in self._crop:
                    self._crop[ck] = torch.zeros((max(rows, 128), 128), dtype=torch.bfloat16, device=self.dev)
                crop = self._crop[ck]
                ops.crop_relu(lc["full"][b * lc["V"]:(b + 1) * lc["V"]], lc["V"], lc["Hs"], lc["Ws"], 128, lc["hf"], lc["wf"],
                              lc["relu_crop"], crop)
            else:
                crop = lc["crop"][b * rows:(b + 1) * rows]
            out.append(localizer_step.SceneContext(
                lp=lp, views=lc["views"][b], fimg=lc["fimg"][b], crop=crop, xs=lc["xs"], ys=lc["ys"], zs=lc["zs"][b],
                plane=plane.features[b].reshape(cells, -1), plane_valid=plane.valid[b].reshape(cells),
                aerial_plane=aerial.features[b].reshape(cells, -1).contiguous() if aerial is not None else None,
                fused_plane=fusedp.features[b].reshape(cells, -1) if fusedp is not None else None,
                top_k=sel.config.top_k_view_selection if select else None,
                view_centers=lc["centers"][b] if select else None,
                max_view_distance=sel.config.get("max_view_distance") if select else None))
        return out

    def loss_and_gradients(self, data: Dict, rngs: Optional[Dict] = None) -> Tuple[Dict, Dict, Dict]:
        """Forward + loss + backward of one batch; the gradients (mean over the batch, `trainer.py:221`) are left in
        `self.bucket` (un-reduced).  Returns (pred, losses, metrics)."""
        m, c = self.model, self.model.config
        if data.get("T_query2map") is None:
            raise ValueError("training needs data['T_query2map'] (the ground truth is the first scored pose)")
        if self.train_encoder:
            data = self._encode_images(data)
        pred = m.apply({"params": self.params}, data, train=False, rngs=rngs)
        losses, metrics = m.loss_metrics_function(pred, data, self.params)
        maps, plane_q, plane_map = pred["similarity_maps"], pred["query"]["bev_matching"], pred["map"]["bev_matching"]
        B, D = plane_map.features.shape[0], plane_q.features.shape[-1]
        remove = c.threshold_remove_accurate_poses
        dr = dt = None
        if remove is not None:
            raise NotImplementedError("threshold_remove_accurate_poses: pass the per-sample errors of loc_nll (not wired)")
        q_xy = torch.from_numpy(np.ascontiguousarray(m.q_xy_p[:, 0])).to(self.dev)
        valid_j = plane_map.valid.contiguous() if c.mask_score_out_of_bounds else None
        g = self.bwd.backward(maps, plane_q.features.reshape(B, -1, D).contiguous(), plane_map.features.contiguous(), q_xy,
                              valid_j, pred["map_t_query_samples"], pred["scores_poses"], m.grid_map.cell_size,
                              c.mask_score_out_of_bounds, c.clip_negative_scores, remove, dr, dt,
                              self._contexts(pred["map"], "bev_mapper"),
                              self._contexts(pred["query"], self.sides[-1]), example_weights=self._example_weights(data, B))
        for v, path in zip(self.bucket.views, self.paths):
            if path == ("temperature",):
                v.copy_(g["temperature"].reshape(1))
                continue
            side, leaf = path[0], path[1:]
            lift, head = (self.bwd.lift_map, self.bwd.head_map) if side == "bev_mapper" else (self.bwd.lift_q, self.bwd.head_q)
            if leaf[0] == "matching_proj":
                v.copy_(head.g[leaf[1]].reshape(v.shape))
            else:
                t = lift.g["/".join(leaf[1:])]
                v.copy_(t[: v.shape[0]].reshape(v.shape) if t.shape != v.shape else t)   # Dense_0 kernel rows are padded
        pred["encoder_cotangents"] = g["encoder_cotangents"]
        if self.train_encoder:
            ec, B = g["encoder_cotangents"], plane_map.features.shape[0]
            lc = pred["map"]["streetview"]["lift_context"]
            Vm, Vq = lc["V"], pred["query"]["streetview"]["lift_context"]["V"]
            self.enc.backward(ec["map"] + ec["query"], [Vm] * B + [Vq] * B, lc["hf"], lc["wf"])
            for v, (_, _, gt) in zip(self.enc_bucket.views, self.enc.leaves):
                v.copy_(gt.reshape(v.shape))
        return pred, losses, metrics

    def _example_weights(self, data: Dict, B: int) -> Optional[torch.Tensor]:
        """`trainer.py:221`: the loss is the mean over `batch['batch_mask']` (padded last batches): weight of example b in
        the mean-over-B gradient the backward kernels produce = mask_b * B / sum(mask).  None without a mask."""
        mask = data.get("batch_mask")
        if mask is None:
            return None
        m = np.asarray(mask, dtype=F).reshape(-1)
        if m.shape[0] != B or m.sum() <= 0:
            raise ValueError("batch_mask must have one entry per example and at least one valid example")
        return torch.from_numpy((m * (B / m.sum())).astype(F)).to(self.dev)

    def _encode_images(self, data: Dict) -> Dict:
        """Training forward of the shared image encoder over the map and the query images of the step; the two pyramids are
        handed to the mappers through `data[...]['image_feature_pyr']` (`StreetViewEncoder.apply` skips its own encoder)."""
        dev = self.dev
        as_dev = lambda x: (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=F))).to(dev)
        im, iq = as_dev(data["map"]["images"]), as_dev(data["query"]["images"])
        B, Vm, H, W, _ = im.shape
        Vq = iq.shape[1]
        if tuple(iq.shape[2:4]) != (H, W):
            raise NotImplementedError("train_encoder: map and query images must have the same size (one encoder batch)")
        enc = self._encoder(B * (Vm + Vq), H, W)
        feats = enc.forward(torch.cat([im.reshape(B * Vm, H, W, 3), iq.reshape(B * Vq, H, W, 3)]))
        hf, wf = -(-H // 4), -(-W // 4)                        # image_encoder.py:137-141: crop to ceil(input / stride)
        out = dict(data)
        out["map"] = dict(data["map"], images=im, image_feature_pyr=enc.pyramid(feats[: B * Vm], hf, wf))
        out["query"] = dict(data["query"], images=iq, image_feature_pyr=enc.pyramid(feats[B * Vm:], hf, wf))
        return out

    def apply_update(self) -> bool:
        """Clip, Adam on the masters, rebuild the parameter tree; the whole update is skipped when a gradient is not
        finite (`trainer.py:260-276` keeps the old parameters AND optimiser state).  Returns whether it was applied."""
        buckets = [self.bucket] + ([self.enc_bucket] if self.enc_bucket is not None else [])
        if not all(bool(b.all_finite().item()) for b in buckets):
            self.skipped_steps += 1
            return False
        if self.max_grad_norm is not None:
            norm = float(torch.sqrt(sum(b.flat.double().square().sum() for b in buckets)))
            for b in buckets:
                b.flat.mul_(clip_scale(norm, self.max_grad_norm))
        self.step += 1
        lr = self.lr(self.step - 1) if callable(self.lr) else self.lr      # optax schedules are evaluated at the step count before the update
        ops.adam_step(self.masters.flat, self.mom[0], self.mom[1], self.bucket.flat, float(lr), self.step)
        if self.enc_bucket is not None:                                     # in place on the tensors the training forward reads
            for (_, p, _), (m1, m2), g in zip(self.enc.leaves, self.enc_mom, self.enc_bucket.views):
                ops.adam_step(p.view(-1), m1.view(-1), m2.view(-1), g.reshape(-1), float(lr), self.step)
        self._rebuild_params()
        return True

    def _rebuild_params(self) -> None:
        host = [v.cpu().numpy().copy() for v in self.masters.views]
        new = dict(self.params)
        for s in self.sides:
            old = self.params[s]
            sv = dict(old["streetview_encoder"])                 # image_encoder keeps its identity (frozen, cached plan)
            sv["proj_mlp"] = {"Dense_0": {}}
            sv["fusion_mlp"] = {"Dense_0": {}, "Dense_1": {}}
            side = dict(old)
            side["streetview_encoder"], side["matching_proj"] = sv, {}
            new[s] = side
        for path, a in zip(self.paths, host):
            if path == ("temperature",):
                new["temperature"] = np.asarray(a.reshape(()), dtype=F)
                continue
            d = new
            for k in path[:-1]:
                d = d[k]
            d[path[-1]] = a
        self.params = new
        self.bwd.lift_map.load_params(new["bev_mapper"]["streetview_encoder"])
        self.bwd.head_map.load_params(new["bev_mapper"]["matching_proj"])
        if len(self.sides) == 2:
            self.bwd.lift_q.load_params(new["bev_mapper_query"]["streetview_encoder"])
            self.bwd.head_q.load_params(new["bev_mapper_query"]["matching_proj"])

    def train_step(self, data: Dict, rngs: Optional[Dict] = None, update: bool = True) -> Tuple[torch.Tensor, Dict, Dict]:
        """`trainer.py:165-295` for one batch.  Returns (per-example total loss, losses, metrics); `metrics['is_finite']`
        tells whether the update was applied."""
        pred, losses, metrics = self.loss_and_gradients(data, rngs)
        self.bucket.allreduce_mean()                                   # jax.lax.pmean(grad, 'batch')
        sq = self.bucket.flat.double().square().sum()
        if self.enc_bucket is not None:
            self.enc_bucket.allreduce_mean()                           # the ~94 MB of encoder gradients: one more collective
            sq = sq + self.enc_bucket.flat.double().square().sum()
        metrics = dict(metrics)
        metrics["l2_grads"] = float(torch.sqrt(sq))
        finite = bool(self.bucket.all_finite().item()) and (self.enc_bucket is None or bool(self.enc_bucket.all_finite().item()))
        metrics["is_finite"] = self.apply_update() if update else finite
        return losses["total"], losses, metrics

    def encoder_params_tree(self) -> Dict:
        """The trained image-encoder parameters as the Flax tree `image_encoder` (host, fp32), for checkpoints."""
        ref = self.params["bev_mapper"]["streetview_encoder"]["image_encoder"]
        out: Dict = {}
        for path, p, _ in self.enc.leaves:
            d = out
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = p.detach().cpu().numpy().copy().reshape(np.asarray(self._get(ref, path)).shape)
        return out

    def grads_tree(self) -> Dict:
        out: Dict = {}
        for path, v in zip(self.paths, self.bucket.views):
            d = out
            for k in path[:-1]:
                d = d.setdefault(k, {})
            d[path[-1]] = v.cpu().numpy().copy().reshape(np.asarray(self._get(self.params, path)).shape)
        return out
