"""B200 implementation behind the `ImageEncoder` module surface of `snap/models/image_encoder.py:97-144`
(BiT ResNet-v2 `snap/models/resnet.py` + FPN decoder), bf16 activations, fp32 accumulation/statistics.

Every conv is one launch of the segmented tcgen05 GEMM (`csrc/gemm_tc.cuh`): 1x1 convs are plain GEMMs
over the NHWC activation matrix, 3x3 convs are 9 row-shifted K-segments over a zero-bordered
(stride 1) or phase-split (stride 2) copy that the GroupNorm-apply kernel writes directly.
The host code below only sequences launches on the current CUDA stream and owns the workspace; it
can be captured into a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _cache, _lib, configs, ops, types


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _WeightBank:
    """Device copies of all conv/dense kernels of one encoder: fp32 masters (Flax HWIO layout) and
    the bf16 [Cout, K] GEMM operands produced by the batched standardisation kernel."""

    def __init__(self, device: torch.device):
        self.device = device
        self.entries: List[Tuple[np.ndarray, int, int, int, bool]] = []  # (w[K,Cout], K, Cout, ldb, std)
        self.b_mats: List[torch.Tensor] = []

    def add(self, kernel: np.ndarray, standardize: bool, k_multiple: int = 64) -> int:
        w = np.ascontiguousarray(kernel, dtype=np.float32)
        cout = w.shape[-1]
        k = int(np.prod(w.shape[:-1]))
        ldb = _round_up(k, k_multiple)
        self.entries.append((w.reshape(k, cout), k, cout, ldb, standardize))
        return len(self.entries) - 1

    def finalize(self) -> None:
        dev = self.device
        total = sum(e[0].size for e in self.entries)
        flat = np.empty(total, np.float32)
        offs, o = [], 0
        for w, *_ in self.entries:
            flat[o:o + w.size] = w.reshape(-1)
            offs.append(o)
            o += w.size
        self.master = torch.from_numpy(flat).to(dev)
        descs = (_lib.WeightDesc * len(self.entries))()
        mapA, mapB = [], []
        self.partials: List[torch.Tensor] = []
        for i, (w, k, cout, ldb, std) in enumerate(self.entries):
            rows = max(cout, 16)
            b = torch.zeros((rows, ldb), dtype=torch.bfloat16, device=dev)
            self.b_mats.append(b)
            ksplit = max(1, min(32, (k + 255) // 256))
            part = torch.zeros((ksplit, cout, 2), dtype=torch.float32, device=dev)
            self.partials.append(part)
            d = descs[i]
            d.w = self.master.data_ptr() + 4 * offs[i]
            d.out = b.data_ptr()
            d.partial = part.data_ptr()
            d.K, d.Cout, d.ldb, d.standardize, d.ksplit = k, cout, ldb, int(std), ksplit
            ct = (cout + 31) // 32
            if std:
                mapA += [(i, c, s, 0) for c in range(ct) for s in range(ksplit)]
            mapB += [(i, c, t, 0) for c in range(ct) for t in range((ldb + 255) // 256)]
        raw = np.frombuffer(bytes(descs), dtype=np.uint8).copy()
        self.descs = torch.from_numpy(raw).to(dev)
        self.mapA = torch.tensor(mapA, dtype=torch.int32, device=dev) if mapA else None
        self.mapB = torch.tensor(mapB, dtype=torch.int32, device=dev)

    def run(self) -> None:
        """StdConv standardisation + relayout of every kernel (2 launches); part of every forward."""
        ops.std_weights_batched(self.descs, self.mapA, self.mapB)


class EncoderPlan:
    """Launch plan + workspace of one ImageEncoder for a fixed input shape [n_img, H, W, 3]."""

    def __init__(self, params: Dict, config, n_img: int, H: int, W: int, device: torch.device,
                 fused_gn="auto"):
        # fused_gn: GroupNorm+ReLU fused into the conv's A-operand path (`snapb200_conv_gn_bf16`, one launch per
        # conv, no normalised copy in HBM).  "auto" (default): see `run_unit`.  "1x1": every stride-1 1x1 conv (conv1, conv3, conv_proj, FPN
        # skip convs) normalises its raw input in shared memory; the 3x3 convs keep the GroupNorm-apply kernel, which
        # writes their zero-bordered / phase-split operand.  True: also the 3x3 convs (slower).  False: no fusion.
        self.fused_gn = fused_gn
        self.halo_conv = os.environ.get("SNAPB200_HALO_CONV", "1") != "0"   # A/B switch for measurements
        enc_cfg = config.encoder
        self.cfg = config
        self.n, self.H, self.W = n_img, H, W
        self.dev = device
        self.skip_root = bool(enc_cfg.skip_root_block)
        width = int(64 * enc_cfg.width)
        blocks = configs.get_block_desc(enc_cfg.depth)
        if enc_cfg.limit_num_blocks is not None:
            blocks = blocks[: enc_cfg.limit_num_blocks]
        self.blocks = list(blocks)
        nlev = config.num_pyr_levels or len(blocks)
        assert nlev == len(blocks), "num_pyr_levels < number of stages is not supported yet"
        self.max_stride = (0 if self.skip_root else 2) + nlev - 1  # image_encoder.py:109-111
        s = 2 ** self.max_stride
        self.Hp, self.Wp = H + (s - H % s), W + (s - W % s)  # pad_to_multiple pads a full stride if aligned
        p_enc, p_dec = params["encoder"], params["decoder"]
        bank = self.bank = _WeightBank(device)
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1)).to(device)
        bf = lambda rows, c: torch.zeros((_round_up(max(rows, 128), 128), c), dtype=torch.bfloat16, device=device)

        # ---- root ----
        if self.skip_root:
            self.root_w = bank.add(p_enc["conv_root"]["kernel"], True, k_multiple=32)
            self.root_geom = (3, 3, 1, 1)
            H0, W0 = self.Hp, self.Wp
        else:
            self.root_w = bank.add(p_enc["root_block"]["conv_root"]["kernel"], True, k_multiple=32)
            self.root_geom = (7, 7, 2, 3)
            H0, W0 = self.Hp // 2, self.Wp // 2
        self.rootHW = (H0, W0)
        self.root_kp = bank.entries[self.root_w][3]
        # implicit root conv: packed bf16 image (+ one row of slack for the last window) and per-kernel-row weights
        kh_, kw_, st_, pd_ = self.root_geom
        self.root_pack = ops.root_packed_geometry(self.Hp, self.Wp, kh_, kw_, st_, pd_)
        cp_, Hq_, Wq_, Ho_, Wo_ = self.root_pack
        assert (Ho_, Wo_) == (H0, W0), (Ho_, Wo_, H0, W0)
        self.img_packed = torch.zeros(n_img * Hq_ * Wq_ * cp_ + 64, dtype=torch.bfloat16, device=device)
        self.root_b = torch.zeros((max(width, 16), kh_ * 32), dtype=torch.bfloat16, device=device)
        self.y_root = bf(n_img * H0 * W0, width)
        if self.skip_root:
            self.x0, self.x0HW = self.y_root, (H0, W0)
        else:
            self.x0HW = (H0 // 2, W0 // 2)
            self.x0 = bf(n_img * self.x0HW[0] * self.x0HW[1], width)

        # ---- stages ----
        self.units = []
        cin = width
        h, w = self.x0HW
        max_rows_c = 0
        self.stage_out = []
        for si, nunits in enumerate(blocks):
            nmid, nout = width * 2 ** si, width * 2 ** si * 4
            for u in range(nunits):
                pu = p_enc[f"block{si + 1}"][f"unit{u + 1:02d}"]
                stride = 2 if (si > 0 and u == 0) else 1
                ho, wo = h // stride, w // stride
                unit = dict(cin=cin, nmid=nmid, nout=nout, stride=stride, h=h, w=w, ho=ho, wo=wo,
                            gn=[(f32(pu[g]["scale"]), f32(pu[g]["bias"])) for g in ("gn1", "gn2", "gn3")],
                            w1=bank.add(pu["conv1"]["kernel"], True), w2=bank.add(pu["conv2"]["kernel"], True),
                            w3=bank.add(pu["conv3"]["kernel"], True),
                            wproj=bank.add(pu["conv_proj"]["kernel"], True) if "conv_proj" in pu else None)
                if stride == 1:
                    unit["a2"] = bf(n_img * (h + 2) * (w + 2), nmid)       # zero-bordered
                else:
                    hq, wq = h // 2 + 1, w // 2 + 1
                    unit["a2"] = bf(4 * n_img * hq * wq, nmid)             # phase-split
                    unit["a1_sub"] = bf(n_img * ho * wo, cin)
                # the (h, w, stride) geometry repeats for units 2.. of a stage: share the bordered buffer
                if u >= 2:
                    unit["a2"] = self.units[-1]["a2"]
                unit["out"] = bf(n_img * ho * wo, nout)
                if u >= 3:
                    unit["out"] = self.units[-2]["out"]  # ping-pong inside the stage
                self.units.append(unit)
                max_rows_c = max(max_rows_c, n_img * h * w * cin, n_img * ho * wo * nout)
                cin, h, w = nout, ho, wo
            # the last unit's output is an FPN skip feature: give it a private buffer
            self.units[-1]["out"] = bf(n_img * h * w, cin)
            self.stage_out.append((self.units[-1]["out"], h, w, cin))
        self.buf_a = torch.zeros(max_rows_c + 128 * 2048, dtype=torch.bfloat16, device=device)  # gn1 / gn3 outputs
        self.buf_y = torch.zeros(max_rows_c + 128 * 2048, dtype=torch.bfloat16, device=device)  # conv1 / conv2 outputs
        self.buf_res = torch.zeros(max_rows_c + 128 * 2048, dtype=torch.bfloat16, device=device)
        # GroupNorm accumulators (sum, sumsq) f64 [slot, replica, img, 32, 2]: 3 per unit (gn1 input, conv1 out, conv2 out),
        # one per FPN level (statistics of relu(stage output)), one scratch; zeroed once per forward.
        self.gn_acc_all = torch.zeros((3 * len(self.units) + len(blocks) + 1, ops.GN_REPLICAS, n_img, 32, 2),
                                      dtype=torch.float64, device=device)
        for i, u in enumerate(self.units):
            u["acc"] = [self.gn_acc_all[3 * i + k] for k in range(3)]
        self.fpn_acc = [self.gn_acc_all[3 * len(self.units) + k] for k in range(len(blocks))]
        self.acc_scratch = self.gn_acc_all[-1]
        ends = np.cumsum(blocks) - 1   # index of the last unit of each stage
        for si, e in enumerate(ends):
            self.units[e]["fpn_acc"] = self.fpn_acc[si]

        # ---- FPN ----
        self.fpn = []
        od = config.output_dim
        for level in range(nlev):
            x, hh, ww, cc = self.stage_out[nlev - 1 - level]
            self.fpn.append(dict(x=x, h=hh, w=ww, c=cc,
                                 gn=(f32(p_dec[f"{level}_skip_norm"]["scale"]), f32(p_dec[f"{level}_skip_norm"]["bias"])),
                                 wk=bank.add(p_dec[f"{level}_skip_conv"]["kernel"], False),
                                 out=bf(n_img * hh * ww, od), up=bf(n_img * hh * ww, od) if level > 0 else None))
        bank.finalize()
        self.out_dim = od
        pad = (self.Hp, self.Wp)
        self.strides = [(pad[0] / f["h"], pad[1] / f["w"]) for f in self.fpn]

    def _view(self, buf: torch.Tensor, rows: int, c: int) -> torch.Tensor:
        return buf[: _round_up(max(rows, 128), 128) * c].view(-1, c)

    def run_root(self, images: torch.Tensor, im2col_done: bool = False) -> torch.Tensor:
        """resnet.py:82-100,199-208.  Also produces the GroupNorm statistics of the first unit's input."""
        n = self.n
        assert tuple(images.shape) == (n, self.H, self.W, 3), images.shape
        kh, kw, st, pd = self.root_geom
        H0, W0 = self.rootHW
        acc0 = self.units[0]["acc"][0]
        cp, Hq, Wq, _, _ = self.root_pack
        width = self.y_root.shape[1]
        if not im2col_done:
            ops.root_pack_image(images, self.Hp, self.Wp, pd, cp, Hq, Wq, self.img_packed)
        ops.root_pack_weights(self.bank.b_mats[self.root_w], width, kh, kw, cp, self.root_b)
        if self.skip_root:
            ops.root_conv(self.img_packed, n, Hq, Wq, cp, kh, st, H0, W0, self.root_b, width, self.y_root, gn_acc=acc0)
        else:
            ops.root_conv(self.img_packed, n, Hq, Wq, cp, kh, st, H0, W0, self.root_b, width, self.y_root)
            ops.maxpool3x3s2(self.y_root, n, H0, W0, self.y_root.shape[1], self.x0)
            ops.gn_stats(self.x0, n, self.x0HW[0] * self.x0HW[1], self.y_root.shape[1], False, acc0)
        return self.x0

    def run_unit(self, u: Dict, x: torch.Tensor, next_acc: Optional[torch.Tensor] = None,
                 forced_input: bool = False) -> torch.Tensor:
        """One pre-activation bottleneck (`snap/models/resnet.py:103-134`); x: bf16 [>= n*h*w, cin].

        The statistics of x must already be in u['acc'][0] (written by the producer's epilogue) unless
        `forced_input` (teacher-forced tests), in which case they are computed here.  The epilogue of conv3
        accumulates the statistics of the unit's output into `next_acc` (next unit's gn1) and, for the last
        unit of a stage, the statistics of relu(output) for the FPN."""
        n, B = self.n, self.bank.b_mats
        cin, nmid, nout, s = u["cin"], u["nmid"], u["nout"], u["stride"]
        h, w, ho, wo = u["h"], u["w"], u["ho"], u["wo"]
        rows_in, rows_out = n * h * w, n * ho * wo
        acc1, acc2, acc3 = u["acc"]
        if forced_input:
            acc1.zero_(); acc2.zero_(); acc3.zero_()
            ops.gn_stats(x, n, h * w, cin, False, acc1)
        if self.fused_gn is True:
            gn1, gn2, gn3 = u["gn"]
            # every conv normalises its input on the fly (GroupNorm + ReLU fused into the GEMM's A producer)
            if u["wproj"] is not None:  # resnet.py:121-122 (projection of the pre-activated tensor, strided)
                res = self._view(self.buf_res, rows_out, nout)
                ops.conv_gn(x, n, h, w, cin, acc1, gn1[0], gn1[1], B[u["wproj"]], res, taps=1, stride=s)
            else:
                res = x
            y1 = self._view(self.buf_y, rows_in, nmid)
            ops.conv_gn(x, n, h, w, cin, acc1, gn1[0], gn1[1], B[u["w1"]], y1, taps=1, stride=1, gn_acc=acc2)
            y2 = self._view(self.buf_a, rows_out, nmid)
            ops.conv_gn(y1, n, h, w, nmid, acc2, gn2[0], gn2[1], B[u["w2"]], y2, taps=9, stride=s, gn_acc=acc3)
            fpn_acc = u.get("fpn_acc")
            if forced_input and fpn_acc is not None:
                fpn_acc.zero_()
            if next_acc is None and fpn_acc is not None:
                next_acc = self.acc_scratch
            ops.conv_gn(y2, n, ho, wo, nmid, acc3, gn3[0], gn3[1], B[u["w3"]], u["out"], taps=1, stride=1,
                        residual=res, gn_acc=next_acc, gn_acc_relu=fpn_acc)
            return u["out"]
        # fused_gn == "1x1": every stride-1 1x1 conv normalises its raw input inside the GEMM (A_TGN1); only the 3x3
        # conv's input goes through the GroupNorm-apply kernel (it writes the zero-bordered / phase-split copy).
        # "auto" (default) fuses where it is measured to pay (profiles/r02_notes.md): conv1 / conv_proj of the units
        # whose bottleneck width fits one N tile (stages 1-2 with 128 x 64 / 128 x 128 tiles: the 352 / 176 MB
        # pre-activation copy disappears; stage 3 with 128 x 256 tiles at one CTA per SM: 64 vs 42 + 36 us) and the FPN
        # skip convs; conv3 (epilogue-bound: residual + statistics) and stage 4 (M = 10,752 rows) keep the apply pass.
        f1 = self.fused_gn == "1x1" or (self.fused_gn == "auto" and nmid <= 256)
        f3 = self.fused_gn == "1x1"
        gn1, gn2, gn3 = u["gn"]
        joined = True
        y1 = self._view(self.buf_y, rows_in, nmid)
        if f1 and s == 1:
            if u["wproj"] is not None:  # resnet.py:121-122 (projection of the pre-activated tensor)
                res = self._view(self.buf_res, rows_out, nout)
                main, side = torch.cuda.current_stream(), self._side()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.conv_gn(x, n, h, w, cin, acc1, gn1[0], gn1[1], B[u["wproj"]], res)
                joined = False
            else:
                res = x
            ops.conv_gn(x, n, h, w, cin, acc1, gn1[0], gn1[1], B[u["w1"]], y1, gn_acc=acc2)
        else:
            a1 = self._view(self.buf_a, rows_in, cin)
            ops.gn_apply(x, n, h, w, cin, acc1, gn1[0], gn1[1], False, True, ops.LAYOUT_DENSE, a1, u.get("a1_sub"))
            if u["wproj"] is not None:
                res = self._view(self.buf_res, rows_out, nout)
                # conv_proj and conv1 both read a1 and are independent: run the projection on a side stream
                main, side = torch.cuda.current_stream(), self._side()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.gemm(u["a1_sub"] if s == 2 else a1, B[u["wproj"]], res, m_rows=rows_out)
                joined = False
            else:
                res = x
            ops.gemm(a1, B[u["w1"]], y1, m_rows=rows_in, gn_acc=acc2, gn_rows_per_img=h * w)
        if s == 1:
            ops.gn_apply(y1, n, h, w, nmid, acc2, u["gn"][1][0], u["gn"][1][1], False, True, ops.LAYOUT_PADDED, u["a2"])
            hp, wp = h + 2, w + 2
            seg = [(a - 1) * wp + (b - 1) for a in range(3) for b in range(3)]
            m_rows, remap = n * hp * wp, (hp, wp, 1, 1, h, w)
        else:
            ops.gn_apply(y1, n, h, w, nmid, acc2, u["gn"][1][0], u["gn"][1][1], False, True, ops.LAYOUT_PHASE, u["a2"])
            hq, wq = h // 2 + 1, w // 2 + 1
            plane = n * hq * wq
            seg = [((a % 2) * 2 + (b % 2)) * plane + (a // 2) * wq + (b // 2) for a in range(3) for b in range(3)]
            m_rows, remap = plane, (hq, wq, 0, 0, ho, wo)
        y2 = self._view(self.buf_y, rows_out, nmid)  # y1 is dead once a2 is written
        if s == 1 and getattr(self, "halo_conv", True) and ops.conv3x3_halo_supported(nmid, nmid, w):
            # stages 1-2: one shared-memory halo block per K chunk serves all nine taps (csrc/conv3x3_halo.cu)
            ops.conv3x3_halo(u["a2"], n, h, w, nmid, B[u["w2"]], y2, gn_acc=acc3)
        else:
            ops.gemm(u["a2"], B[u["w2"]], y2, m_rows=m_rows, seg_off=seg, seg_k=nmid, remap=remap,
                     gn_acc=acc3, gn_rows_per_img=ho * wo)
        if not joined:
            # a3 reuses the buffer of a1, which the projection (side stream) is still reading; conv3 adds its result
            torch.cuda.current_stream().wait_stream(self._side())
        fpn_acc = u.get("fpn_acc")
        if forced_input and fpn_acc is not None:
            fpn_acc.zero_()
        if next_acc is None and fpn_acc is not None:
            next_acc = self.acc_scratch
        if f3:
            # conv2 wrote y2 (raw): conv3 normalises it while loading
            ops.conv_gn(y2, n, ho, wo, nmid, acc3, gn3[0], gn3[1], B[u["w3"]], u["out"], residual=res,
                        gn_acc=next_acc, gn_acc_relu=fpn_acc)
            return u["out"]
        a3 = self._view(self.buf_a, rows_out, nmid)
        ops.gn_apply(y2, n, ho, wo, nmid, acc3, gn3[0], gn3[1], False, True, ops.LAYOUT_DENSE, a3)
        ops.gemm(a3, B[u["w3"]], u["out"], m_rows=rows_out, residual=res, gn_acc=next_acc, gn_acc_relu=fpn_acc,
                 gn_rows_per_img=ho * wo)
        return u["out"]

    def run_fpn(self, forced_input: bool = False) -> List[torch.Tensor]:
        """`snap/models/image_encoder.py:79-94` over the stage outputs (coarse -> fine)."""
        n, B = self.n, self.bank.b_mats
        outs, prev = [], None
        nlev = len(self.fpn)
        for level, f in enumerate(self.fpn):
            rows = n * f["h"] * f["w"]
            acc = self.fpn_acc[nlev - 1 - level]
            if forced_input:
                acc.zero_()
                ops.gn_stats(f["x"], n, f["h"] * f["w"], f["c"], True, acc)
            if self.fused_gn:  # True, "1x1" or "auto": the skip conv is a 1x1 conv
                if prev is not None:
                    ops.upsample2x(prev["out"], n, prev["h"], prev["w"], self.out_dim, f["up"])
                # relu -> GroupNorm -> 1x1 conv (+ up-sampled coarser level) in one launch
                ops.conv_gn(f["x"], n, f["h"], f["w"], f["c"], acc, f["gn"][0], f["gn"][1], B[f["wk"]], f["out"],
                            taps=1, stride=1, pre_relu=True, post_relu=False,
                            residual=f["up"] if prev is not None else None)
            else:
                a = self._view(self.buf_a, rows, f["c"])
                ops.gn_apply(f["x"], n, f["h"], f["w"], f["c"], acc, f["gn"][0], f["gn"][1], True, False,
                             ops.LAYOUT_DENSE, a)
                if prev is not None:
                    ops.upsample2x(prev["out"], n, prev["h"], prev["w"], self.out_dim, f["up"])
                ops.gemm(a, B[f["wk"]], f["out"], m_rows=rows, residual=f["up"] if prev is not None else None)
            outs.append(f["out"][:rows].view(n, f["h"], f["w"], self.out_dim))
            prev = f
        return outs

    def _side(self) -> torch.cuda.Stream:
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.dev)
        return self._side_stream

    def run(self, images: torch.Tensor) -> List[torch.Tensor]:
        """images f32 [n,H,W,3] in [0,1] (device) -> FPN features coarse->fine, each [n,h,w,C] bf16 (uncropped)."""
        main, side = torch.cuda.current_stream(), self._side()
        # the StdConv weight standardisation does not depend on the images: overlap it with the image packing
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.bank.run()
        self.gn_acc_all.zero_()
        kh, kw, st, pd = self.root_geom
        cp, Hq, Wq, _, _ = self.root_pack
        ops.root_pack_image(images, self.Hp, self.Wp, pd, cp, Hq, Wq, self.img_packed)
        main.wait_stream(side)
        x = self.run_root(images, im2col_done=True)
        for i, u in enumerate(self.units):
            nxt = self.units[i + 1]["acc"][0] if i + 1 < len(self.units) else None
            x = self.run_unit(u, x, nxt)
        return self.run_fpn()

    def cropped_shapes(self) -> List[Tuple[int, int]]:
        """image_encoder.py:137-141: crop each level to ceil(input / stride)."""
        return [(int(round(math.ceil(self.H / s[0]))), int(round(math.ceil(self.W / s[1])))) for s in self.strides]


class ImageEncoder:
    """Mirror of `snap.models.image_encoder.ImageEncoder` (config, dtype) -> FeatureImagePyramid.

    `apply(variables, image, train=False)` follows `flax.linen.Module.apply`; the launch plan for an
    input shape is built on first use and cached."""

    default_config = staticmethod(configs.image_encoder)

    def __init__(self, config=None, dtype=torch.bfloat16, fused_gn="auto"):
        self.fused_gn = fused_gn
        if dtype != torch.bfloat16:
            raise NotImplementedError("the B200 path computes in bf16 (fp32 accumulate / statistics)")
        self.config = config if config is not None else configs.image_encoder()
        if self.config.encoder_name != "resnet":
            raise ValueError(self.config.encoder_name)  # image_encoder.py:112-113
        self.dtype = dtype
        self._plans = _cache.ParamCache()

    def plan(self, params: Dict, n: int, H: int, W: int, device) -> EncoderPlan:
        return self._plans.lookup(params, (n, H, W, str(device)),
                                  lambda: EncoderPlan(params, self.config, n, H, W, device, fused_gn=self.fused_gn))

    def clear_cache(self) -> None:
        self._plans.clear()

    def apply(self, variables: Dict, image: torch.Tensor, train: bool = False) -> types.FeatureImagePyramid:
        if train:
            raise NotImplementedError("training (backward kernels) is a 'next' row of SURVEY.md §8(f)")
        if not image.is_cuda:
            raise _lib.SnapB200Error("ImageEncoder needs CUDA tensors: there is no CPU fallback")
        params = variables["params"] if "params" in variables else variables
        n, H, W, _ = image.shape
        plan = self.plan(params, n, H, W, image.device)
        feats = plan.run(image.float().contiguous())
        crops = plan.cropped_shapes()
        return types.FeatureImagePyramid(
            features=[f[:, :h, :w, :] for f, (h, w) in zip(feats, crops)],
            strides=[np.asarray(s) for s in plan.strides], uncropped=feats)

    __call__ = apply
