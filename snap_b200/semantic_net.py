"""B200 implementation of the semantic head forward of `snap.models.semantic_net.SemanticNet`
(`snap/models/semantic_net.py:145-198`, decoder_type='resnet_stage'): Dense(128->dim) -> ResNetStage(num_units)
-> MLP(dim -> dim -> num_classes) on the BEV plane, logits f32, zero where the plane is invalid.

The residual units reuse the encoder's launch plan (`image_encoder.EncoderPlan.run_unit`): the BEV plane is an
'image' of G x G pixels."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import configs, image_encoder, ops, types

F = np.float32


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
        if self.config.decoder_type != "resnet_stage":
            raise NotImplementedError("only decoder_type='resnet_stage' (snap/configs/train_semantics.py:28-30) is built")
        c = self.config
        self.num_area = len(c.area_classes)
        self.num_excl = len(c.object_classes_exclusive)
        self.num_classes = self.num_area + (self.num_excl + len(c.object_classes_independent) + 1
                                            if (c.object_classes_exclusive or c.object_classes_independent) else 0)
        self._cache: Dict = {}

    def _plan(self, params: Dict, B: int, G0: int, G1: int, C: int, dev):
        key = (id(params), B, G0, G1, C, str(dev))
        if key not in self._cache:
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
            self._cache[key] = dict(stage=stage, w=w, x0=z(dim), h=z(dim), logits=z(16))
        return self._cache[key]

    def apply(self, variables: Dict, plane: types.FeaturePlane) -> Dict:
        params = variables["params"] if "params" in variables else variables
        params = params.get("decoder", params)
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
