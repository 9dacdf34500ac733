"""B200 implementation of the semantic head forward of `snap.models.semantic_net.SemanticNet`
(`snap/models/semantic_net.py:145-198`, decoder_type='resnet_stage'): Dense(128->dim) -> ResNetStage(num_units)
-> MLP(dim -> dim -> num_classes) on the BEV plane, logits f32, zero where the plane is invalid.

The residual units reuse the encoder's launch plan (`image_encoder.EncoderPlan.run_unit`): the BEV plane is an
'image' of G x G pixels."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import _cache, configs, image_encoder, ops, types

F = np.float32


def apply_batch_mask(dlogits: torch.Tensor, data, B: int) -> None:
    """`trainer.py:221`: the reference means the loss over `batch['batch_mask']`; the loss-gradient kernel divides by B.
    Rescale example b's rows of dlogits (bf16 [B * cells, C]) by mask_b * B / sum(mask) in place (no-op without a mask;
    the factor is exact for the usual all-ones mask and one bf16 rounding otherwise)."""
    mask = data.get("batch_mask") if isinstance(data, dict) else None
    if mask is None:
        return
    m = np.asarray(mask, dtype=np.float32).reshape(-1)
    if m.shape[0] != B or m.sum() <= 0:
        raise ValueError("batch_mask must have one entry per example and at least one valid example")
    w = torch.from_numpy((m * (B / m.sum())).astype(np.float32)).to(dlogits.device)
    v = dlogits[: B * (dlogits.shape[0] // B)].view(B, -1, dlogits.shape[1])
    v.copy_((v.float() * w.view(B, 1, 1)).to(dlogits.dtype))


class _StagePlan(image_encoder.EncoderPlan):
    """A ResNetStage(block_size, nmid=C/4, stride 1) over [n_img, H, W, C] with the EncoderPlan unit machinery."""

    def __init__(self, stage_params: Dict, n_img: int, H: int, W: int, C: int, device, fused_gn: bool = False):  # noqa
        self.fused_gn, self.n, self.dev = fused_gn, n_img, device
        self._side_stream = None
        bank = self.bank = image_encoder._WeightBank(device)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1)).to(device)
        bf = lambda rows, c: torch.zeros((image_encoder._round_up(max(rows, 128), 128), c), dtype=torch.bfloat16, device=device)
        self.units = []
        nmid, nout = C // 4, C
        names = sorted(k for k in stage_params if k.startswith("unit"))
        for name in names:
            pu = stage_params[name]
            u = dict(cin=C, nmid=nmid, nout=nout, stride=1, h=H, w=W, ho=H, wo=W,
                     gn=[(f32(pu[g]["scale"]), f32(pu[g]["bias"])) for g in ("gn1", "gn2", "gn3")],
                     w1=bank.add(pu["conv1"]["kernel"], True), w2=bank.add(pu["conv2"]["kernel"], True),
                     w3=bank.add(pu["conv3"]["kernel"], True), wproj=None,
                     a2=bf(n_img * (H + 2) * (W + 2), nmid), out=bf(n_img * H * W, nout))
            self.units.append(u)
        self.extra = {}
        rows_c = n_img * H * W * C
        self.buf_a = torch.zeros(rows_c + 128 * 2048, dtype=torch.bfloat16, device=device)
        self.buf_y = torch.zeros(rows_c + 128 * 2048, dtype=torch.bfloat16, device=device)
        self.buf_res = self.buf_y
        self.gn_acc_all = torch.zeros((3 * len(self.units) + 1, ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64,
                                      device=device)
        for i, u in enumerate(self.units):
            u["acc"] = [self.gn_acc_all[3 * i + k] for k in range(3)]
        self.acc_scratch = self.gn_acc_all[-1]


class SemanticHead:
    """decoder of `SemanticNet` (`semantic_net.py:145-165,185-197`).  `apply(variables, plane)` takes the
    `bev_features` FeaturePlane of `BEVMapper` and returns the reference's logits dict."""

    def __init__(self, config=None, dtype=torch.bfloat16):
        self.config = config if config is not None else configs.semantic_net()
        if self.config.decoder_type not in ("resnet_stage", "mlp"):
            raise ValueError(f"Unknown {self.config.decoder_type}")                           # semantic_net.py:164-165
        c = self.config
        self.num_area = len(c.area_classes)
        self.num_excl = len(c.object_classes_exclusive)
        self.num_classes = self.num_area + (self.num_excl + len(c.object_classes_independent) + 1
                                            if (c.object_classes_exclusive or c.object_classes_independent) else 0)
        self._cache = _cache.ParamCache()

    def clear_cache(self) -> None:
        self._cache.clear()

    def _plan(self, params: Dict, B: int, G0: int, G1: int, C: int, dev):
        def build():
            dim = self.config.decoder_dim
            stage = _StagePlan(params["layers_1"], B, G0, G1, dim, dev)
            bank = stage.bank
            f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1)).to(dev)
            k1 = np.zeros((dim, 16), F)
            k1[:, : self.num_classes] = params["layers_3"]["Dense_1"]["kernel"]
            b1 = np.zeros((16,), F)
            b1[: self.num_classes] = params["layers_3"]["Dense_1"]["bias"]
            w = dict(d0=bank.add(params["layers_0"]["kernel"], False), m0=bank.add(params["layers_3"]["Dense_0"]["kernel"], False),
                     m1=bank.add(k1, False), d0_b=f32(params["layers_0"]["bias"]),
                     m0_b=f32(params["layers_3"]["Dense_0"]["bias"]), m1_b=f32(b1))
            bank.finalize()
            rows = image_encoder._round_up(max(B * G0 * G1, 128), 128)
            z = lambda c, dt=torch.bfloat16: torch.zeros((rows, c), dtype=dt, device=dev)
            return dict(stage=stage, w=w, x0=z(dim), h=z(dim), logits=z(16))
        return self._cache.lookup(params, (B, G0, G1, C, str(dev)), build)

    def apply(self, variables: Dict, plane: types.FeaturePlane) -> Dict:
        params = variables["params"] if "params" in variables else variables
        params = params.get("decoder", params)
        if self.config.decoder_type == "mlp":                                                # :147-152
            return self._cache.lookup(params, ("mlp", str(plane.features.device)),
                                      lambda: MLPHeadTrainer(self.config, params, plane.features.device)).forward(plane)
        f, valid = plane.features.contiguous(), plane.valid.contiguous()
        B, G0, G1, C = f.shape
        dev = f.device
        pl = self._plan(params, B, G0, G1, C, dev)
        stage, w = pl["stage"], pl["w"]
        Bm = stage.bank.b_mats
        rows = B * G0 * G1
        stage.bank.run()
        stage.gn_acc_all.zero_()
        dim = self.config.decoder_dim
        ops.gemm(f.view(rows, C), Bm[w["d0"]], pl["x0"], m_rows=rows, bias=w["d0_b"])          # nn.Dense (:154-158)
        x = pl["x0"]
        ops.gn_stats(x, B, G0 * G1, dim, False, stage.units[0]["acc"][0])
        for i, u in enumerate(stage.units):                                                      # ResNetStage (:159)
            nxt = stage.units[i + 1]["acc"][0] if i + 1 < len(stage.units) else None
            x = stage.run_unit(u, x, nxt)
        ops.gemm(x, Bm[w["m0"]], pl["h"], m_rows=rows, bias=w["m0_b"], relu=True)                # MLP Dense_0 + relu
        ops.gemm(pl["h"], Bm[w["m1"]], pl["logits"], m_rows=rows, bias=w["m1_b"],
                 row_mask=valid.view(rows))                                                      # Dense_1, zero invalid (:186)
        logits = pl["logits"][:rows, : self.num_classes].float().view(B, G0, G1, self.num_classes)
        pred = {"logits_areas": logits[..., : self.num_area]}
        if self.num_classes > self.num_area:                                                     # :190-196
            rest = logits[..., self.num_area:]
            pred["logits_objects_exclusive"] = rest[..., : self.num_excl + 1]
            pred["logits_objects_independent"] = rest[..., self.num_excl + 1:]
        return pred

    __call__ = apply



class MLPHeadTrainer:
    """Forward and head-only training step of the default 'mlp' decoder (`semantic_net.py:147-152`: layers.MLP with
    layers (dim,) * mlp_num_layers + (num_classes,)) on frozen BEV features, as `snap/configs/train_semantics.py:35-36`
    fine-tunes it (`freeze_params_reg_exp = 'bev_mapper/'`): forward -> loss (`:300-343`) -> mean over the batch
    (`trainer.py:221`) -> backward through the MLP -> gradient mean over ranks (`trainer.py:231-234`) -> optax.adam
    (`defaults.py:81`).  Master parameters and Adam moments are fp32 device arrays (the reference keeps them in the
    model dtype); GEMM operands are re-derived from the masters at every step."""

    def __init__(self, config, params: Dict, device, lr: float = 5e-5):
        c = self.config = config
        self.dev = device
        self.lr = lr
        self.num_area = len(c.area_classes)
        self.num_excl = len(c.object_classes_exclusive) + 1 if (c.object_classes_exclusive or c.object_classes_independent) else 0
        self.num_indep = len(c.object_classes_independent)
        self.num_classes = self.num_area + self.num_excl + self.num_indep
        names = [f"Dense_{i}" for i in range(c.mlp_num_layers + 1)]
        if sorted(params) != names:
            raise KeyError(f"decoder params must be {names} (layers.MLP), got {sorted(params)}")
        self.names = names
        bank = self.bank = image_encoder._WeightBank(device)
        self.dims = []
        for n in names:
            k = np.ascontiguousarray(params[n]["kernel"], dtype=F)
            cin, cout = k.shape
            last = n == names[-1]
            if last:                       # logits padded to 32 columns (GEMM N multiple of 16, backward K multiple of 32)
                kp = np.zeros((cin, 32), F)
                kp[:, :cout] = k
                k = kp
            if cin % 32 or k.shape[1] % 16:
                raise NotImplementedError("MLP widths must be multiples of 32")
            bank.add(k, False, k_multiple=32)
            self.dims.append((cin, k.shape[1], cout))
        bank.finalize()
        # views of the fp32 masters [in, out(padded)] inside the bank + biases, gradients, Adam moments
        self.W, o = [], 0
        for cin, coutp, _ in self.dims:
            self.W.append(bank.master[o:o + cin * coutp].view(cin, coutp))
            o += cin * coutp
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        self.b = []
        for n, (cin, coutp, cout) in zip(names, self.dims):
            b = z(coutp)
            b[:cout] = torch.from_numpy(np.ascontiguousarray(params[n]["bias"], dtype=F)).to(device)
            self.b.append(b)
        from . import parallel
        # gradients live in one flat fp32 bucket: the mean over ranks is a single in-place all-reduce (parallel.GradBucket)
        self.bucket = parallel.GradBucket([(cin, coutp) for cin, coutp, _ in self.dims] + [(coutp,) for _, coutp, _ in self.dims],
                                          device)
        self.dW = self.bucket.views[: len(self.dims)]
        self.db = self.bucket.views[len(self.dims):]
        self.mom = [[z(*t.shape), z(*t.shape)] for t in self.W + self.b]
        self.Wc = [z(cin, coutp, dt=torch.bfloat16) for cin, coutp, _ in self.dims]   # bf16 [in, out]: B operand of dX
        self.step = 0
        self._buf: Dict = {}

    def _buffers(self, rows: int):
        if rows not in self._buf:
            R = image_encoder._round_up(max(rows, 128), 128)
            z = lambda c, dt=torch.bfloat16: torch.zeros((R, c), dtype=dt, device=self.dev)
            self._buf[rows] = dict(h=[z(coutp) for _, coutp, _ in self.dims], d=[z(cin) for cin, _, _ in self.dims],
                                   dlogits=z(32), counts=torch.zeros((1, 2), dtype=torch.float32, device=self.dev))
        return self._buf[rows]

    def forward(self, plane: types.FeaturePlane, keep: bool = False) -> Dict:
        f, valid = plane.features.contiguous(), plane.valid.contiguous()
        B, G0, G1, C = f.shape
        rows = B * G0 * G1
        buf = self._buffers(rows)
        self.bank.run()
        x = f.view(rows, C)
        last = len(self.names) - 1
        for i in range(len(self.names)):                                   # Dense -> (relu -> Dense)* (layers.py:73-77)
            ops.gemm(x, self.bank.b_mats[i], buf["h"][i], m_rows=rows, bias=self.b[i], relu=i < last,
                     row_mask=valid.view(rows) if i == last else None)     # zero invalid cells (:186)
            x = buf["h"][i]
        logits = buf["h"][last][:rows, : self.num_classes].float().view(B, G0, G1, self.num_classes)
        pred = {"logits_areas": logits[..., : self.num_area]}
        if self.num_classes > self.num_area:
            rest = logits[..., self.num_area:]
            pred["logits_objects_exclusive"] = rest[..., : self.num_excl]
            pred["logits_objects_independent"] = rest[..., self.num_excl:]
        return pred

    def params_tree(self) -> Dict:
        """Current parameters as the Flax tree (host, fp32)."""
        return {n: {"kernel": self.W[i][:, :cout].cpu().numpy().copy(), "bias": self.b[i][:cout].cpu().numpy().copy()}
                for i, (n, (_, _, cout)) in enumerate(zip(self.names, self.dims))}

    def train_step(self, plane: types.FeaturePlane, model: "SemanticNetModel", data: Dict, update: bool = True):
        """One training step on frozen BEV features.  Returns (loss tensor [], losses, metrics); gradients stay in
        `self.dW` / `self.db` (averaged over ranks when torch.distributed is initialised)."""
        from . import parallel
        pred = self.forward(plane)
        pred["bev_features"] = plane
        losses, metrics, ctx = model.loss_metrics_function(pred, data, return_context=True)
        B, cells = ctx["logits"].shape[:2]
        rows = B * cells
        buf = self._buffers(rows)
        if buf["counts"].shape[0] != B:
            buf["counts"] = torch.zeros((B, 2), dtype=torch.float32, device=self.dev)
        ops.sem_loss_grad(ctx["logits"], ctx["labels_area"], ctx["valid_area"], ctx["labels_excl"], ctx["masks_indep"],
                          ctx["valid"], self.num_area, self.num_excl, self.num_indep, ctx["weights"], buf["counts"],
                          buf["dlogits"])
        apply_batch_mask(buf["dlogits"], data, B)
        x0 = plane.features.contiguous().view(rows, -1)
        acts = [x0] + buf["h"][:-1]                       # input of layer i
        dy = buf["dlogits"]
        for i in reversed(range(len(self.names))):
            cin, coutp, _ = self.dims[i]
            ops.dense_wgrad(acts[i], dy, rows, cin, coutp, self.dW[i], self.db[i])
            if i > 0:                                     # dX = dY W^T, then the ReLU in front of this layer
                ops.cast_pad_bf16(self.W[i], self.Wc[i])
                ops.gemm(dy, self.Wc[i], buf["d"][i], m_rows=rows, seg_k=coutp)
                ops.relu_bwd(acts[i], buf["d"][i], rows * cin)
                dy = buf["d"][i]
        self.bucket.allreduce_mean()                      # jax.lax.pmean(grad, 'batch') (trainer.py:231-234)
        if update and bool(self.bucket.all_finite().item()):   # non-finite gradients: keep parameters and Adam state (trainer.py:260-276)
            self.step += 1
            for k, (pt, g) in enumerate(zip(self.W + self.b, list(self.dW) + list(self.db))):
                ops.adam_step(pt.view(-1) if pt.is_contiguous() else pt, self.mom[k][0].view(-1), self.mom[k][1].view(-1),
                              g.view(-1), self.lr, self.step)
        return losses["total"], losses, metrics


def balancing_weights(frequencies: Dict[str, float], classes, binary: bool = False, eps: float = 1e-3):
    """`semantic_net.py:31-53` (host side: a handful of class weights)."""
    f = np.array([frequencies[c] for c in classes], dtype=np.float64)
    if not binary:
        f = f / f.sum()
    f = f.clip(min=eps)
    w = (1 / (f * len(classes))).astype(F)
    if binary:
        return w, (1 / ((1 - f).clip(min=eps) * len(classes))).astype(F)
    return w


def create_exclusive_labels(masks_all: np.ndarray, gt_classes, classes, add_void: bool = False):
    """`SemanticNetModel._create_exclusive_labels` (`semantic_net.py:254-276`), host side (label preparation)."""
    gi = {c: i for i, c in enumerate(gt_classes)}
    masks = masks_all[..., [gi[c] for c in classes]].copy()
    if "line" in classes:
        ml = masks_all[..., gi["line"]].copy()
        for c in ("stopline", "otherlanemarking"):
            if c in gi and c not in classes:
                ml |= masks_all[..., gi[c]]
        masks[..., list(classes).index("line")] = ml
    valid = masks.any(-1)
    labels = np.argmax(masks, -1)
    if add_void:
        labels = np.where(valid, labels, len(classes))
    return labels.astype(np.int32), valid


class SemanticNetModel:
    """Loss / metric side of `snap.models.semantic_net.SemanticNetModel` (`semantic_net.py:300-343`) for the logits of
    `SemanticHead`; `gt_classes` = dataset_meta_data['semantic_classes_gt']."""

    def __init__(self, config=None, gt_classes=()):
        self.config = config if config is not None else configs.semantic_net()
        self.gt_classes = tuple(gt_classes)
        self._weights: Dict = {}

    def loss_metrics_function(self, pred: Dict, data: Dict, model_params=None, return_context: bool = False):
        """pred: logits dict of SemanticHead + 'bev_features' (FeaturePlane); data['rasters']['gt_semantics'] bool
        [B,G,G,N_gt] (NumPy).  Returns (losses, metrics) as dicts of per-example device tensors."""
        c = self.config
        if "map" in data:
            data = data["map"]
        logits_a = pred["logits_areas"]
        dev = logits_a.device
        B, G0, G1, Ka = logits_a.shape
        cells = G0 * G1
        masks = data["rasters"]["gt_semantics"]                      # bool / u8 [B,G,G,N_gt]: NumPy or a CUDA tensor
        if not isinstance(masks, torch.Tensor):
            masks = torch.from_numpy(np.ascontiguousarray(masks).view(np.uint8) if masks.dtype == bool
                                     else np.ascontiguousarray(masks, dtype=np.uint8))
        masks = masks.to(dev).to(torch.uint8).reshape(B * cells, -1).contiguous()
        ngt = masks.shape[1]
        assert ngt == len(self.gt_classes), "gt_semantics channels must follow semantic_classes_gt"
        gi = {n: i for i, n in enumerate(self.gt_classes)}

        def selection(classes):                                      # _create_exclusive_labels (:254-270)
            sel = []
            for n in classes:
                idx = [gi[n]]
                if n == "line":
                    idx += [gi[x] for x in ("stopline", "otherlanemarking") if x in gi and x not in classes]
                sel.append(idx)
            return sel
        bev_valid = pred["bev_features"].valid.reshape(B, cells).contiguous()
        has_obj = "logits_objects_exclusive" in pred
        parts = [logits_a]
        Ke = Ki = 0
        le_d = mi_d = None
        la_d = torch.empty((B, cells), dtype=torch.int32, device=dev)
        valid_area = torch.empty((B, cells), dtype=torch.uint8, device=dev)
        sel_e, sel_i = [], []
        if has_obj:
            parts += [pred["logits_objects_exclusive"], pred["logits_objects_independent"]]
            Ke, Ki = parts[1].shape[-1], parts[2].shape[-1]
            sel_e = selection(c.object_classes_exclusive)            # + void (:278-281)
            sel_i = [gi[n] for n in c.object_classes_independent]
            le_d = torch.empty((B, cells), dtype=torch.int32, device=dev)
            mi_d = torch.empty((B, cells, Ki), dtype=torch.uint8, device=dev)
        ops.sem_labels(selection(c.area_classes), sel_e, sel_i, ngt, masks, bev_valid, la_d, valid_area, le_d, mi_d)  # :301-302
        logits = torch.cat(parts, -1).reshape(B, cells, Ka + Ke + Ki).contiguous()          # [areas | excl | indep]
        weights = None
        fa, fo = c.get("area_frequencies"), c.get("object_frequencies")
        if fa or fo:
            wkey = (str(dev), has_obj)
            if wkey not in self._weights:    # class-balancing weights (:31-53), uploaded once per device
                t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(dev)
                w = [None] * 4
                if fa:
                    w[0] = t(balancing_weights(dict(fa), c.area_classes))
                if fo and has_obj:
                    w[1] = t(balancing_weights(dict(fo), (*c.object_classes_exclusive, "void")))
                    wp, wn = balancing_weights(dict(fo), c.object_classes_independent, binary=True)
                    w[2], w[3] = t(wp), t(wn)
                self._weights[wkey] = w
            weights = self._weights[wkey]
        out = torch.empty((B, ops._lib.SEM_OUT), dtype=torch.float32, device=dev)
        ops.sem_loss(logits, la_d, valid_area, le_d, mi_d, bev_valid, Ka, Ke, Ki, weights, out)
        losses = {"nll_areas": out[:, 0], "total": out[:, 3]}
        metrics = {"accuracy": out[:, 4], "recall/average": out[:, 6]}
        for i, n in enumerate(c.area_classes):
            metrics[f"recall/{n}"] = out[:, 16 + i]
        if has_obj:
            losses.update(nll_objects_exclusive=out[:, 1], nll_objects_indep=out[:, 2])
            metrics.update({"accuracy/excl": out[:, 5], "recall/average/excl": out[:, 7], "recall/average/indep": out[:, 8]})
            for i, n in enumerate((*c.object_classes_exclusive, "void")):
                metrics[f"recall/{n}"] = out[:, 24 + i]
            for i, n in enumerate(c.object_classes_independent):
                metrics[f"recall/{n}"] = out[:, 32 + i]
        metrics = {f"semantics/{k}": v for k, v in metrics.items()}
        if return_context:   # device-side labels / masks / weights, reused by the backward of the training step
            return losses, metrics, dict(logits=logits, labels_area=la_d, valid_area=valid_area,
                                         labels_excl=le_d, masks_indep=mi_d, valid=bev_valid, weights=weights)
        return losses, metrics
