"""Backward of the FPN decoder of the image encoder (`snap/models/image_encoder.py:53-94`, norm='bit_resnet': per level
ReLU -> GroupNorm -> 1x1 conv (no bias) -> + x2 bilinear up-sampling of the coarser level): the first block of the encoder
backward.  Only the finest level is consumed downstream (`streetview_encoder.py:222`), so its cotangent
enters at the last level and walks towards the coarse ones through the up-sampling.

Per level, fine -> coarse (tensors as the forward plan `image_encoder.EncoderPlan.run_fpn` holds them):

    a   = GroupNorm(relu(x))                        recomputed (`snapb200_gn_apply`, pre_relu)
    dW  = a^T dout                                  `snapb200_dense_wgrad` over <= 1024-channel slices of a
    da  = dout W^T                                  GEMM engine, B = the transposed kernel (`snapb200_wt_segments`)
    dx  = GroupNorm backward, masked by x > 0       `snapb200_gn_backward(pre_relu=1, post_relu=0)`  (+ dscale, dbias)
    dout(coarser) = up-sampling backward of dout    `snapb200_upsample2x_backward`

The cotangents dx of the skip inputs (the stage outputs of the ResNet trunk) are returned: the trunk's own backward
(strided bottleneck units, root block) is not built yet."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import image_encoder, ops

F = np.float32


class FPNBackward:
    def __init__(self, decoder_params: Dict, n_img: int, shapes: List, device, out_dim: int = 128):
        """decoder_params: the Flax tree `image_encoder/decoder` ('{level}_skip_norm', '{level}_skip_conv'); shapes: per
        level (coarse -> fine) the (h, w, c) of the skip input."""
        self.n, self.dev, self.od = n_img, device, out_dim
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1).copy()).to(device)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        bank = self.bank = image_encoder._WeightBank(device)
        self.lv = []
        for level, (h, w, c) in enumerate(shapes):
            rows = n_img * h * w
            if rows % 16:
                raise NotImplementedError("n * h * w must be a multiple of 16 at every level (split-K weight-gradient kernel)")
            R = image_encoder._round_up(max(rows, 128), 128)
            self.lv.append(dict(h=h, w=w, c=c, rows=rows,
                                scale=f32(decoder_params[f"{level}_skip_norm"]["scale"]),
                                bias=f32(decoder_params[f"{level}_skip_norm"]["bias"]),
                                wk=bank.add(decoder_params[f"{level}_skip_conv"]["kernel"], False),
                                a=z(R, c, dt=torch.bfloat16), da=z(R, c, dt=torch.bfloat16), dx=z(R, c, dt=torch.bfloat16),
                                bt=z(c, out_dim, dt=torch.bfloat16), dup=z(R, out_dim, dt=torch.bfloat16),
                                out=z(R, out_dim, dt=torch.bfloat16), up=z(R, out_dim, dt=torch.bfloat16),
                                accb=z(n_img, c, 2, dt=torch.float64),
                                g=dict(kernel=z(c, out_dim), scale=z(c), bias=z(c))))
        bank.finalize()

    def forward(self, skips: List[torch.Tensor], accs: List[torch.Tensor]) -> torch.Tensor:
        """The FPN forward on this object's buffers (`EncoderPlan.run_fpn`); returns the finest level bf16 [n*h*w, out_dim]."""
        n, od, Bm = self.n, self.od, self.bank.b_mats
        self.bank.run()
        prev = None
        for level, L in enumerate(self.lv):
            ops.gn_apply(skips[level], n, L["h"], L["w"], L["c"], accs[level], L["scale"], L["bias"], True, False,
                         ops.LAYOUT_DENSE, L["a"])
            if prev is not None:
                ops.upsample2x(prev["out"], n, prev["h"], prev["w"], od, L["up"])
            ops.gemm(L["a"], Bm[L["wk"]], L["out"], m_rows=L["rows"], residual=L["up"] if prev is not None else None)
            prev = L
        return prev["out"]

    def backward(self, skips: List[torch.Tensor], accs: List[torch.Tensor], dout_finest: torch.Tensor) -> List[torch.Tensor]:
        """skips[level] bf16 [n*h*w, c] (coarse -> fine) = the FPN inputs; accs[level] = the f64 GroupNorm accumulators of
        relu(skip) the forward used; dout_finest bf16 [n*h*w, out_dim] of the LAST level.  Fills the per-level gradients
        (`grads_tree`) and returns the cotangents of the skip inputs, coarse -> fine."""
        n, od = self.n, self.od
        self.bank.run()
        dout = dout_finest
        dskips = [None] * len(self.lv)
        for level in reversed(range(len(self.lv))):
            L = self.lv[level]
            h, w, c, rows = L["h"], L["w"], L["c"], L["rows"]
            ops.gn_apply(skips[level], n, h, w, c, accs[level], L["scale"], L["bias"], True, False, ops.LAYOUT_DENSE, L["a"])
            for c0 in range(0, c, 1024):                       # dW = a^T dout, <= 1024 input channels per launch
                kc = min(1024, c - c0)
                ops.dense_wgrad(L["a"][:, c0:c0 + kc], dout, rows, kc, od, L["g"]["kernel"][c0:c0 + kc], None)
            ops.wt_segments(self.bank.b_mats[L["wk"]], od, c, 1, L["bt"])
            ops.gemm(dout, L["bt"], L["da"], m_rows=rows, seg_k=od)
            ops.gn_backward(skips[level], L["da"], n, h, w, c, accs[level], L["scale"], L["bias"], L["accb"], L["dx"],
                            L["g"]["scale"], L["g"]["bias"], post_relu=False, pre_relu=True)
            dskips[level] = L["dx"]
            if level > 0:                                       # out = conv + upsample2x(coarser out): both get dout
                P = self.lv[level - 1]
                ops.upsample2x_backward(dout, n, P["h"], P["w"], od, P["dup"])
                dout = P["dup"]
        return dskips

    def grads_tree(self) -> Dict:
        out: Dict = {}
        for level, L in enumerate(self.lv):
            out[f"{level}_skip_norm"] = {"scale": L["g"]["scale"].cpu().numpy().reshape(1, 1, 1, -1).copy(),
                                         "bias": L["g"]["bias"].cpu().numpy().reshape(1, 1, 1, -1).copy()}
            out[f"{level}_skip_conv"] = {"kernel": L["g"]["kernel"].cpu().numpy().reshape(1, 1, L["c"], self.od).copy()}
        return out


class BottleneckUnitTrainer:
    """Training forward + backward of ONE stride-1 pre-activation bottleneck unit of the ResNet trunk
    (`snap/models/resnet.py:103-134`), with the identity shortcut (cin == nout) or the 1x1 projection of the
    PRE-ACTIVATED input (`:121-122`; first unit of stage 1), for any of the trunk's widths (GroupNorm backward up to 2048
    channels; weight gradients in <= 1024-wide slices).  The launch plan is the one of `semantic_train.StageHeadTrainer`
    (3x3 dX = mirrored 9-segment GEMM over the zero-bordered cotangent, dW = nine row-shifted split-K products), checked
    on the emulated operator layer against autograd of the oracle's `residual_unit`.  Stride-2 units (phase-split layout)
    are the next step of the encoder backward."""

    def __init__(self, unit_params: Dict, n_img: int, H: int, W: int, device):
        p = unit_params
        k1, k2, k3 = (np.ascontiguousarray(p[c]["kernel"], dtype=F) for c in ("conv1", "conv2", "conv3"))
        self.cin, self.nmid, self.nout = k1.shape[2], k1.shape[3], k3.shape[3]
        self.proj = "conv_proj" in p
        if not self.proj and self.cin != self.nout:
            raise ValueError("a unit without conv_proj needs cin == nout (identity shortcut)")
        self.n, self.H, self.W, self.dev = n_img, H, W, device
        rows = self.rows = n_img * H * W
        if rows % 16:
            raise NotImplementedError("n * H * W must be a multiple of 16 (split-K weight-gradient kernel)")
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1).copy()).to(device)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        bf = lambda r, c: z(r, c, dt=torch.bfloat16)
        bank = self.bank = image_encoder._WeightBank(device)
        self.w = {c: bank.add(p[c]["kernel"], True) for c in ("conv1", "conv2", "conv3") + (("conv_proj",) if self.proj else ())}
        bank.finalize()
        off, self.master = 0, {}
        for name, (_, k, cout, _, _) in zip(self.w, bank.entries):
            self.master[name] = bank.master[off: off + k * cout].view(k, cout)
            off += k * cout
        self.gn = {g: (f32(p[g]["scale"]), f32(p[g]["bias"])) for g in ("gn1", "gn2", "gn3")}
        self.g = {name: z(*m.shape) for name, m in self.master.items()}
        self.gs = {name: z(*m.shape) for name, m in self.master.items()}
        self.ggn = {g: (z(c), z(c)) for g, c in (("gn1", self.cin), ("gn2", self.nmid), ("gn3", self.nmid))}
        R = image_encoder._round_up(max(rows, 128), 128)
        self.Mb = n_img * (H + 2) * (W + 2)
        Rb = image_encoder._round_up(self.Mb + 64, 128)
        cin, nmid, nout = self.cin, self.nmid, self.nout
        self.acc = torch.zeros((3, ops.GN_REPLICAS, n_img, 32, 2), dtype=torch.float64, device=device)
        self.b = dict(a1=bf(R, cin), y1=bf(R, nmid), a2=bf(Rb, nmid), y2=bf(R, nmid), a3=bf(R, nmid), out=bf(R, nout),
                      res=bf(R, nout) if self.proj else None, da3=bf(R, nmid), dc2b=bf(Rb, nmid), da2=bf(R, nmid),
                      dc1=bf(R, nmid), da1p=bf(R, cin) if self.proj else None, da1=bf(R, cin), dx=bf(R, cin),
                      accb=z(n_img, max(cin, nmid), 2, dt=torch.float64), tmp=z(1024, 1024))
        self.bt = dict(conv1=bf(cin, nmid), conv2=bf(nmid, 9 * nmid), conv3=bf(nmid, nout),
                       conv_proj=bf(cin, nout) if self.proj else None)

    def _wgrad(self, x: torch.Tensor, dy: torch.Tensor, M: int, K: int, N: int, out: torch.Tensor) -> None:
        """out f32 [K, N] = x^T dy with the split-K kernel's K, N <= 1024 limit lifted by slicing."""
        for k0 in range(0, K, 1024):
            kc = min(1024, K - k0)
            for n0 in range(0, N, 1024):
                nc = min(1024, N - n0)
                if nc == N:
                    ops.dense_wgrad(x[:, k0:k0 + kc], dy, M, kc, N, out[k0:k0 + kc], None)
                else:           # column slices of dW are not contiguous: through a scratch tile
                    tmp = self.b["tmp"].view(-1)[: kc * nc].view(kc, nc)
                    ops.dense_wgrad(x[:, k0:k0 + kc], dy[:, n0:n0 + nc], M, kc, nc, tmp, None)
                    out[k0:k0 + kc, n0:n0 + nc].copy_(tmp)

    def forward(self, x: torch.Tensor, acc_x: torch.Tensor, next_acc=None) -> torch.Tensor:
        """x bf16 [>= n*H*W, cin]; acc_x: the GroupNorm accumulators of x (f64 [REPLICAS, n, 32, 2])."""
        n, H, W, rows, Bm, b = self.n, self.H, self.W, self.rows, self.bank.b_mats, self.b
        hp, wp = H + 2, W + 2
        seg = [(a - 1) * wp + (c - 1) for a in range(3) for c in range(3)]
        self.bank.run()
        self.acc.zero_()
        self.x, self.acc_x = x, acc_x
        ops.gn_apply(x, n, H, W, self.cin, acc_x, *self.gn["gn1"], False, True, ops.LAYOUT_DENSE, b["a1"])
        res = x
        if self.proj:                                                                        # resnet.py:121-122
            ops.gemm(b["a1"], Bm[self.w["conv_proj"]], b["res"], m_rows=rows)
            res = b["res"]
        ops.gemm(b["a1"], Bm[self.w["conv1"]], b["y1"], m_rows=rows, gn_acc=self.acc[0], gn_rows_per_img=H * W)
        ops.gn_apply(b["y1"], n, H, W, self.nmid, self.acc[0], *self.gn["gn2"], False, True, ops.LAYOUT_PADDED, b["a2"])
        ops.gemm(b["a2"], Bm[self.w["conv2"]], b["y2"], m_rows=self.Mb, seg_off=seg, seg_k=self.nmid,
                 remap=(hp, wp, 1, 1, H, W), gn_acc=self.acc[1], gn_rows_per_img=H * W)
        ops.gn_apply(b["y2"], n, H, W, self.nmid, self.acc[1], *self.gn["gn3"], False, True, ops.LAYOUT_DENSE, b["a3"])
        ops.gemm(b["a3"], Bm[self.w["conv3"]], b["out"], m_rows=rows, residual=res, gn_acc=next_acc, gn_rows_per_img=H * W)
        return b["out"]

    def backward(self, dout: torch.Tensor) -> torch.Tensor:
        """dout bf16 [>= rows, nout] -> dx bf16 [rows.., cin]; parameter gradients in `grads_tree()`."""
        n, H, W, rows, Bm, b = self.n, self.H, self.W, self.rows, self.bank.b_mats, self.b
        cin, nmid, nout = self.cin, self.nmid, self.nout
        hp, wp = H + 2, W + 2
        seg = [(a - 1) * wp + (c - 1) for a in range(3) for c in range(3)]
        lo = wp + 1
        Mp = image_encoder._round_up(self.Mb - 2 * lo, 16)
        # conv3
        self._wgrad(b["a3"], dout, rows, nmid, nout, self.gs["conv3"])
        ops.wt_segments(Bm[self.w["conv3"]], nout, nmid, 1, self.bt["conv3"])
        ops.gemm(dout, self.bt["conv3"], b["da3"], m_rows=rows, seg_k=nout)
        ops.gn_backward(b["y2"], b["da3"], n, H, W, nmid, self.acc[1], *self.gn["gn3"], b["accb"], b["dc2b"],
                        *self.ggn["gn3"], post_relu=True, padded_out=True)
        # conv2
        for t, off in enumerate(seg):
            ops.dense_wgrad(b["a2"][lo + off: lo + off + Mp], b["dc2b"][lo: lo + Mp], Mp, nmid, nmid,
                            self.gs["conv2"][t * nmid: (t + 1) * nmid], None)
        ops.wt_segments(Bm[self.w["conv2"]], nmid, nmid, 9, self.bt["conv2"])
        ops.gemm(b["dc2b"], self.bt["conv2"], b["da2"], m_rows=self.Mb, seg_off=[-o for o in seg], seg_k=nmid,
                 remap=(hp, wp, 1, 1, H, W))
        ops.gn_backward(b["y1"], b["da2"], n, H, W, nmid, self.acc[0], *self.gn["gn2"], b["accb"], b["dc1"],
                        *self.ggn["gn2"], post_relu=True)
        # conv1 (+ conv_proj: both read the pre-activated a1)
        self._wgrad(b["a1"], b["dc1"], rows, cin, nmid, self.gs["conv1"])
        ops.wt_segments(Bm[self.w["conv1"]], nmid, cin, 1, self.bt["conv1"])
        add = None
        if self.proj:
            self._wgrad(b["a1"], dout, rows, cin, nout, self.gs["conv_proj"])
            ops.wt_segments(Bm[self.w["conv_proj"]], nout, cin, 1, self.bt["conv_proj"])
            ops.gemm(dout, self.bt["conv_proj"], b["da1p"], m_rows=rows, seg_k=nout)
            add = b["da1p"]
        ops.gemm(b["dc1"], self.bt["conv1"], b["da1"], m_rows=rows, seg_k=nmid, residual=add)
        ops.gn_backward(self.x, b["da1"], n, H, W, cin, self.acc_x, *self.gn["gn1"], b["accb"], b["dx"],
                        *self.ggn["gn1"], post_relu=True, add=None if self.proj else dout)      # identity shortcut (:134)
        for name in self.master:
            ops.stdconv_backward(self.master[name], self.gs[name], self.g[name])
        return b["dx"]

    def grads_tree(self, unit_params: Dict) -> Dict:
        out: Dict = {}
        for name, g in self.g.items():
            out[name] = {"kernel": g.cpu().numpy().reshape(np.asarray(unit_params[name]["kernel"]).shape).copy()}
        for gname, (gs, gb) in self.ggn.items():
            out[gname] = {"scale": gs.cpu().numpy().reshape(1, 1, 1, -1).copy(), "bias": gb.cpu().numpy().reshape(1, 1, 1, -1).copy()}
        return out


class StridedUnitTrainer(BottleneckUnitTrainer):
    """The FIRST unit of stages 2-4 (`resnet.py:103-134` with strides=(2, 2)): the 3x3 conv and the 1x1 projection of the
    pre-activated input subsample by two.  The forward is `EncoderPlan.run_unit`'s stride-2 branch (3x3 conv input in the
    PHASE-SPLIT layout: four H/2+1 x W/2+1 planes of the zero-bordered tensor, nine constant-offset K-segments).  Backward:

        conv2 dW   nine split-K products  a2_phase[plane(tap) + off(tap) + m]^T dc2[m]  with dc2 in the output GEMM's row
                   indexing ([n, H/2+1, W/2+1], bottom / right zero-extended: `snapb200_gn_backward`, out_layout 2)
        conv2 dX   one GEMM per phase plane over the taps of that parity (4 / 2 / 2 / 1 K-segments, mirrored offsets),
                   written in the phase-split layout, which the GroupNorm backward in front reads directly (dy_phase)
        conv_proj  dW from the even-pixel subsample a1_sub; its dX is added at the even pixels inside the GroupNorm
                   backward of the unit's input (dy_sub)."""

    def __init__(self, unit_params: Dict, n_img: int, H: int, W: int, device):
        if "conv_proj" not in unit_params:
            raise ValueError("a stride-2 unit always has conv_proj (resnet.py:117-122)")
        if H % 2 or W % 2:
            raise ValueError("stride-2 units need even H, W")
        super().__init__(unit_params, n_img, H, W, device)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        bf = lambda r, c: z(r, c, dt=torch.bfloat16)
        n, cin, nmid, nout = n_img, self.cin, self.nmid, self.nout
        self.ho, self.wo, self.hq, self.wq = H // 2, W // 2, H // 2 + 1, W // 2 + 1
        self.rows_out = n * self.ho * self.wo
        self.plane = n * self.hq * self.wq
        if self.rows_out % 16:
            raise NotImplementedError("n * (H/2) * (W/2) must be a multiple of 16 (split-K weight-gradient kernel)")
        Ro = image_encoder._round_up(max(self.rows_out, 128), 128)
        self.Mp = image_encoder._round_up(self.plane, 16)
        Rp = image_encoder._round_up(4 * self.plane + self.wq + 1 + self.Mp + 16, 128)
        Rq = image_encoder._round_up(self.Mp + self.wq + 17, 128)
        self.b.update(a1s=bf(Ro, cin), a2=bf(Rp, nmid), y2=bf(Ro, nmid), a3=bf(Ro, nmid), out=bf(Ro, nout), res=bf(Ro, nout),
                      da3=bf(Ro, nmid), dc2b=bf(Rq, nmid), da2=bf(Rp, nmid), da1p=bf(Ro, cin))
        self.btp = [bf(nmid, k * nmid) for k in (4, 2, 2, 1)]        # phase (0,0) (0,1) (1,0) (1,1): taps of that parity
        self.taps = [(a, c) for a in range(3) for c in range(3)]

    def _seg(self):
        return [((a % 2) * 2 + (c % 2)) * self.plane + (a // 2) * self.wq + (c // 2) for a, c in self.taps]

    def forward(self, x: torch.Tensor, acc_x: torch.Tensor, next_acc=None) -> torch.Tensor:
        n, H, W, rows, Bm, b = self.n, self.H, self.W, self.rows, self.bank.b_mats, self.b
        ho, wo, hq, wq = self.ho, self.wo, self.hq, self.wq
        self.bank.run()
        self.acc.zero_()
        self.x, self.acc_x = x, acc_x
        ops.gn_apply(x, n, H, W, self.cin, acc_x, *self.gn["gn1"], False, True, ops.LAYOUT_DENSE, b["a1"], b["a1s"])
        ops.gemm(b["a1s"], Bm[self.w["conv_proj"]], b["res"], m_rows=self.rows_out)              # resnet.py:121-122
        ops.gemm(b["a1"], Bm[self.w["conv1"]], b["y1"], m_rows=rows, gn_acc=self.acc[0], gn_rows_per_img=H * W)
        ops.gn_apply(b["y1"], n, H, W, self.nmid, self.acc[0], *self.gn["gn2"], False, True, ops.LAYOUT_PHASE, b["a2"])
        ops.gemm(b["a2"], Bm[self.w["conv2"]], b["y2"], m_rows=self.plane, seg_off=self._seg(), seg_k=self.nmid,
                 remap=(hq, wq, 0, 0, ho, wo), gn_acc=self.acc[1], gn_rows_per_img=ho * wo)
        ops.gn_apply(b["y2"], n, ho, wo, self.nmid, self.acc[1], *self.gn["gn3"], False, True, ops.LAYOUT_DENSE, b["a3"])
        ops.gemm(b["a3"], Bm[self.w["conv3"]], b["out"], m_rows=self.rows_out, residual=b["res"], gn_acc=next_acc,
                 gn_rows_per_img=ho * wo)
        return b["out"]

    def backward(self, dout: torch.Tensor, add: torch.Tensor = None) -> torch.Tensor:
        """`add` (bf16 dense [n*H*W, cin], optional): a second cotangent of the unit's input, e.g. the FPN's cotangent of the
        previous stage's output, folded into the last GroupNorm backward."""
        n, H, W, rows, Bm, b = self.n, self.H, self.W, self.rows, self.bank.b_mats, self.b
        cin, nmid, nout = self.cin, self.nmid, self.nout
        ho, wo, wq, plane, ro, Mp = self.ho, self.wo, self.wq, self.plane, self.rows_out, self.Mp
        # conv3 on the subsampled grid
        self._wgrad(b["a3"], dout, ro, nmid, nout, self.gs["conv3"])
        ops.wt_segments(Bm[self.w["conv3"]], nout, nmid, 1, self.bt["conv3"])
        ops.gemm(dout, self.bt["conv3"], b["da3"], m_rows=ro, seg_k=nout)
        ops.gn_backward(b["y2"], b["da3"], n, ho, wo, nmid, self.acc[1], *self.gn["gn3"], b["accb"], b["dc2b"],
                        *self.ggn["gn3"], post_relu=True, out_layout=2)
        # conv2 (3x3, stride 2): dW per tap, dX per phase plane
        w2 = Bm[self.w["conv2"]]
        slot = [0, 0, 0, 0]
        offs = [[], [], [], []]
        for t, (a, c) in enumerate(self.taps):
            pl, off = (a % 2) * 2 + (c % 2), (a // 2) * wq + (c // 2)
            ops.dense_wgrad(b["a2"][pl * plane + off: pl * plane + off + Mp], b["dc2b"][:Mp], Mp, nmid, nmid,
                            self.gs["conv2"][t * nmid: (t + 1) * nmid], None)
            ops.wt_segments(w2[:, t * nmid: (t + 1) * nmid], nmid, nmid, 1,
                            self.btp[pl][:, slot[pl] * nmid: (slot[pl] + 1) * nmid])
            slot[pl] += 1
            offs[pl].append(-off)
        for pl in range(4):
            ops.gemm(b["dc2b"], self.btp[pl], b["da2"][pl * plane: (pl + 1) * plane], m_rows=plane, seg_off=offs[pl],
                     seg_k=nmid)
        ops.gn_backward(b["y1"], b["da2"], n, H, W, nmid, self.acc[0], *self.gn["gn2"], b["accb"], b["dc1"],
                        *self.ggn["gn2"], post_relu=True, dy_phase=True)
        # conv1 and the strided projection
        self._wgrad(b["a1"], b["dc1"], rows, cin, nmid, self.gs["conv1"])
        ops.wt_segments(Bm[self.w["conv1"]], nmid, cin, 1, self.bt["conv1"])
        ops.gemm(b["dc1"], self.bt["conv1"], b["da1"], m_rows=rows, seg_k=nmid)
        self._wgrad(b["a1s"], dout, ro, cin, nout, self.gs["conv_proj"])
        ops.wt_segments(Bm[self.w["conv_proj"]], nout, cin, 1, self.bt["conv_proj"])
        ops.gemm(dout, self.bt["conv_proj"], b["da1p"], m_rows=ro, seg_k=nout)
        ops.gn_backward(self.x, b["da1"], n, H, W, cin, self.acc_x, *self.gn["gn1"], b["accb"], b["dx"],
                        *self.ggn["gn1"], post_relu=True, dy_sub=b["da1p"], add=add)
        for name in self.master:
            ops.stdconv_backward(self.master[name], self.gs[name], self.g[name])
        return b["dx"]


class TrunkTrainer:
    """Training forward + backward of the whole image encoder (`resnet.py:184-216` + `image_encoder.py:53-94`) for an input
    whose sides are already multiples of the maximum stride: root block (7x7 / 2 StdConv as im2col GEMM, 3x3 / 2 max pool),
    the four stages of bottleneck units (`BottleneckUnitTrainer` / `StridedUnitTrainer`) and the FPN (`FPNBackward`).  The
    stage outputs feed both the next stage and the FPN: the FPN's cotangent is folded into the next stage's first GroupNorm
    backward (`add`).  The first convolution needs no dX (the image is data): its weight gradient is one split-K product
    with the im2col matrix."""

    def __init__(self, enc_params: Dict, n_img: int, H: int, W: int, device, out_dim: int = 128,
                 skip_root_block: bool = False):
        """skip_root_block: the aerial encoder (`defaults.py:187`, `resnet.py:200-208`): a 3x3 / stride-1 `conv_root` and no
        max pool, i.e. the trunk runs at the raster's resolution and the maximum stride is 8."""
        ms = 8 if skip_root_block else 32
        if H % ms or W % ms:
            raise ValueError(f"H, W must be multiples of {ms} (pad_to_multiple, image_encoder.py:32-39, is the caller's job)")
        pe, pd = enc_params["encoder"], enc_params["decoder"]
        self.n, self.H, self.W, self.dev, self.skip_root = n_img, H, W, device, skip_root_block
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)
        bf = lambda r, c: z(r, c, dt=torch.bfloat16)
        kroot = np.ascontiguousarray((pe if skip_root_block else pe["root_block"])["conv_root"]["kernel"], dtype=F)
        self.root_shape, self.c0 = kroot.shape, kroot.shape[-1]
        self.bank = image_encoder._WeightBank(device)
        self.root_w = self.bank.add(kroot, True, 32)
        self.bank.finalize()
        self.K0 = int(np.prod(kroot.shape[:3]))
        self.Kp = image_encoder._round_up(self.K0, 32)
        self.root_master = self.bank.master[: self.K0 * self.c0].view(self.K0, self.c0)
        h1, w1 = (H, W) if skip_root_block else (H // 2, W // 2)
        h, w = (h1, w1) if skip_root_block else (h1 // 2, w1 // 2)
        self.rows0, self.hw_root, self.hw0 = n_img * h1 * w1, (h1, w1), (h, w)
        R0 = image_encoder._round_up(max(self.rows0, 128), 128)
        self.A = bf(R0, self.Kp)
        self.y_root, self.dy_root = bf(R0, self.c0), bf(R0, self.c0)
        self.x0 = self.y_root if skip_root_block else bf(image_encoder._round_up(max(n_img * h * w, 128), 128), self.c0)
        self.g_root_s, self.g_root = z(self.Kp, self.c0), z(self.K0, self.c0)
        self.units, shapes = [], []
        i = 1
        while f"block{i}" in pe:
            names = sorted(k for k in pe[f"block{i}"] if k.startswith("unit"))
            for j, name in enumerate(names):
                pu = pe[f"block{i}"][name]
                if j == 0 and i > 1:
                    u = StridedUnitTrainer(pu, n_img, h, w, device)
                    h, w = h // 2, w // 2
                else:
                    u = BottleneckUnitTrainer(pu, n_img, h, w, device)
                self.units.append(dict(t=u, path=(f"block{i}", name), params=pu, last=(j == len(names) - 1), hw=(h, w)))
            shapes.append((h, w, self.units[-1]["t"].nout))
            i += 1
        self.acc_in = [z(ops.GN_REPLICAS, n_img, 32, 2, dt=torch.float64) for _ in self.units]
        self.acc_fpn = [z(ops.GN_REPLICAS, n_img, 32, 2, dt=torch.float64) for _ in shapes]
        self.fpn = FPNBackward(pd, n_img, shapes[::-1], device, out_dim)        # coarse -> fine
        self.dec_params = pd

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """images f32 [n, H, W, 3] in [0, 1] -> finest FPN level bf16 [n * H/4 * W/4, out_dim]."""
        n = self.n
        kh, kw = self.root_shape[:2]
        self.bank.run()
        ops.root_im2col(images, self.H, self.W, kh, kw, 1 if self.skip_root else 2, kh // 2, self.A)
        ops.gemm(self.A, self.bank.b_mats[self.root_w], self.y_root, m_rows=self.rows0, seg_k=self.Kp)
        if not self.skip_root:
            ops.maxpool3x3s2(self.y_root, n, self.hw_root[0], self.hw_root[1], self.c0, self.x0)
        for a in self.acc_in + self.acc_fpn:
            a.zero_()
        ops.gn_stats(self.x0, n, self.hw0[0] * self.hw0[1], self.c0, False, self.acc_in[0])
        x, skips = self.x0, []
        for k, u in enumerate(self.units):
            nxt = self.acc_in[k + 1] if k + 1 < len(self.units) else None
            x = u["t"].forward(x, self.acc_in[k], nxt)
            if u["last"]:
                ops.gn_stats(x, n, u["hw"][0] * u["hw"][1], u["t"].nout, True, self.acc_fpn[len(skips)])
                skips.append(x)
        self.skips = skips[::-1]                                                  # coarse -> fine
        self.accs = self.acc_fpn[: len(skips)][::-1]
        return self.fpn.forward(self.skips, self.accs)

    def backward(self, dout_finest: torch.Tensor) -> None:
        n = self.n
        dsk = self.fpn.backward(self.skips, self.accs, dout_finest)[::-1]         # per stage, fine (stage 1) -> coarse
        stage = len(dsk) - 1
        dout = dsk[stage]
        for u in reversed(self.units):
            t = u["t"]
            if isinstance(t, StridedUnitTrainer):
                stage -= 1
                dout = t.backward(dout, add=dsk[stage])                           # + the FPN's cotangent of the stage output
            else:
                dout = t.backward(dout)
        # root block: max pool, then the weight gradient of the first convolution
        if self.skip_root:
            dy_root = dout
        else:
            ops.maxpool3x3s2_backward(self.y_root, dout, n, self.hw_root[0], self.hw_root[1], self.c0, self.dy_root)
            dy_root = self.dy_root
        ops.dense_wgrad(self.A, dy_root, self.rows0, self.Kp, self.c0, self.g_root_s, None)
        ops.stdconv_backward(self.root_master, self.g_root_s[: self.K0], self.g_root)

    def grads_tree(self) -> Dict:
        groot = {"conv_root": {"kernel": self.g_root.cpu().numpy().reshape(self.root_shape).copy()}}
        enc: Dict = dict(groot) if self.skip_root else {"root_block": groot}
        for u in self.units:
            enc.setdefault(u["path"][0], {})[u["path"][1]] = u["t"].grads_tree(u["params"])
        return {"encoder": enc, "decoder": self.fpn.grads_tree()}
