"""Head-only training step of BASELINE configs[4] with the decoder the reference fine-tunes
(`snap/configs/train_semantics.py:27-36`: decoder_type='resnet_stage', dim 256, 2 units, `bev_mapper/` frozen):

    Dense(128 -> dim) -> ResNetStage(num_units) -> MLP(dim -> dim -> num_classes)        (`semantic_net.py:153-161`)

forward (activations kept) -> loss (`:300-343`) -> backward -> gradient mean over ranks (`trainer.py:231-234`) -> Adam.

Launch plan of the backward (closed forms and decomposition checked against torch autograd on the CPU,
`tools/design/backward_formulas.py::residual_unit_backward`):

    dense layers   dW = X^T dY: `snapb200_dense_wgrad`; dX = dY W^T: the tcgen05 GEMM engine, B = the kernel [in, out]
    1x1 StdConv    the same, with B = the transposed standardised kernel (`snapb200_wt_segments`, taps = 1)
    3x3 StdConv    dX: the engine's 9-segment mode over the zero-bordered dY with mirrored row offsets and
                   B = [in, 9*out]; dW: nine row-shifted `snapb200_dense_wgrad` products on the bordered layouts
    GroupNorm+ReLU `snapb200_gn_backward` (ReLU mask recomputed with the forward's rounding chain, one reduction per
                   (image, channel), dx written dense / zero-bordered, identity shortcut added in the same pass)
    StdConv        `snapb200_stdconv_backward` per kernel

Master parameters, gradients and Adam moments are fp32 device arrays; every GEMM operand is re-derived from the masters
at each step (`_WeightBank.run`).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import image_encoder, ops, semantic_net, types

F = np.float32


class StageHeadTrainer:
    def __init__(self, config, params: Dict, device, lr: float = 5e-5):
        c = self.config = config
        if c.decoder_type != "resnet_stage":
            raise ValueError("StageHeadTrainer trains the 'resnet_stage' decoder; use MLPHeadTrainer for 'mlp'")
        params = params.get("decoder", params)
        self.dev, self.lr = device, lr
        self.num_area = len(c.area_classes)
        self.num_excl = len(c.object_classes_exclusive) + 1 if (c.object_classes_exclusive or c.object_classes_independent) else 0
        self.num_indep = len(c.object_classes_independent)
        self.num_classes = self.num_area + self.num_excl + self.num_indep
        dim = self.dim = int(c.decoder_dim)
        nmid = self.nmid = dim // 4
        if dim not in (256,) or self.num_classes > 32:
            raise NotImplementedError("decoder_dim must be 256 (GroupNorm backward kernels: C in {64, 128, 256})")
        unit_names = sorted(k for k in params["layers_1"] if k.startswith("unit"))
        self.cin = int(np.asarray(params["layers_0"]["kernel"]).shape[0])
        f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=F).reshape(-1).copy()).to(device)
        z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=device)

        bank = self.bank = image_encoder._WeightBank(device)
        self.rec: List[Dict] = []     # every trainable array: name path, Flax shape, master p, gradient g
        def add_kernel(path, kernel, std):
            k = np.ascontiguousarray(kernel, dtype=F)
            idx = bank.add(k, std)
            self.rec.append(dict(path=path, shape=k.shape, bank=idx, std=std))
            return idx
        def add_vec(path, vec, pad_to=None):
            v = np.ascontiguousarray(vec, dtype=F).reshape(-1)
            t = z(pad_to or len(v))
            t[: len(v)] = torch.from_numpy(v.copy()).to(device)
            self.rec.append(dict(path=path, shape=np.asarray(vec).shape, p=t, g=z(*t.shape), n=len(v)))
            return t
        self.d0 = add_kernel(("layers_0", "kernel"), params["layers_0"]["kernel"], False)
        self.d0_b = add_vec(("layers_0", "bias"), params["layers_0"]["bias"])
        self.units = []
        for name in unit_names:
            pu = params["layers_1"][name]
            u = dict(name=name)
            for g, cc in (("gn1", dim), ("gn2", nmid), ("gn3", nmid)):
                u[g] = (add_vec(("layers_1", name, g, "scale"), pu[g]["scale"]),
                        add_vec(("layers_1", name, g, "bias"), pu[g]["bias"]))
                u["d" + g + "_idx"] = (len(self.rec) - 2, len(self.rec) - 1)
            for cv in ("conv1", "conv2", "conv3"):
                u[cv] = add_kernel(("layers_1", name, cv, "kernel"), pu[cv]["kernel"], True)
            self.units.append(u)
        self.m0 = add_kernel(("layers_3", "Dense_0", "kernel"), params["layers_3"]["Dense_0"]["kernel"], False)
        self.m0_b = add_vec(("layers_3", "Dense_0", "bias"), params["layers_3"]["Dense_0"]["bias"])
        k1 = np.zeros((dim, 32), F)      # logits padded to 32 columns (GEMM N multiple of 16, backward K multiple of 32)
        k1[:, : self.num_classes] = params["layers_3"]["Dense_1"]["kernel"]
        self.m1 = add_kernel(("layers_3", "Dense_1", "kernel"), k1, False)
        self.rec[-1]["shape"] = (dim, self.num_classes)
        self.m1_b = add_vec(("layers_3", "Dense_1", "bias"), params["layers_3"]["Dense_1"]["bias"], pad_to=32)
        bank.finalize()
        # fp32 master views [K, Cout] inside the bank's flat buffer, in `add` order
        off = 0
        views = []
        for w, k, cout, _, _ in bank.entries:
            views.append(bank.master[off: off + k * cout].view(k, cout))
            off += k * cout
        for r in self.rec:
            if "bank" in r:
                r["p"] = views[r["bank"]]
                r["g"] = z(*r["p"].shape)
                if r["std"]:
                    r["gs"] = z(*r["p"].shape)          # gradient w.r.t. the standardised kernel
        # all gradients live in ONE flat fp32 bucket (parallel.GradBucket): the backward kernels write into its views and
        # the gradient mean over ranks (trainer.py:231-234) is a single in-place all-reduce
        from . import parallel
        self.bucket = parallel.GradBucket([r["g"].shape for r in self.rec], device)
        for r, v in zip(self.rec, self.bucket.views):
            r["g"] = v
        for u in self.units:
            for g in ("gn1", "gn2", "gn3"):
                i, j = u["d" + g + "_idx"]
                u["d" + g] = (self.rec[i]["g"], self.rec[j]["g"])
        self.by_bank = {r["bank"]: r for r in self.rec if "bank" in r}
        self.mom = [[z(*r["p"].shape), z(*r["p"].shape)] for r in self.rec]
        # B operands of the dX GEMMs
        bf = lambda *s: z(*s, dt=torch.bfloat16)
        for u in self.units:
            u["bt1"], u["bt2"], u["bt3"] = bf(dim, nmid), bf(nmid, 9 * nmid), bf(nmid, dim)
        self.wc_m0, self.wc_m1 = bf(dim, dim), bf(dim, 32)
        self.step = 0
        self._buf: Dict = {}

    # ------------------------------------------------------------------------------------------------------------
    def _buffers(self, n: int, H: int, W: int) -> Dict:
        key = (n, H, W)
        if key not in self._buf:
            dev, dim, nmid = self.dev, self.dim, self.nmid
            rows = n * H * W
            if rows % 16:
                raise NotImplementedError("B * G * G must be a multiple of 16 (split-K weight-gradient kernel)")
            R = image_encoder._round_up(max(rows, 128), 128)
            Mb = n * (H + 2) * (W + 2)
            Rb = image_encoder._round_up(Mb + 64, 128)     # slack rows (zero) for the row-shifted weight-gradient products
            bf = lambda r, c: torch.zeros((r, c), dtype=torch.bfloat16, device=dev)
            acc = torch.zeros((3 * len(self.units) + 1, ops.GN_REPLICAS, n, 32, 2), dtype=torch.float64, device=dev)
            units = [dict(a1=bf(R, dim), y1=bf(R, nmid), a2=bf(Rb, nmid), y2=bf(R, nmid), a3=bf(R, nmid), out=bf(R, dim),
                          acc=[acc[3 * i + k] for k in range(3)]) for i in range(len(self.units))]
            need = 0
            f = ops._lib.lib().snapb200_dense_wgrad_workspace
            f.restype = ops.C.c_size_t
            Mp = image_encoder._round_up(Mb - 2 * (W + 3), 16)
            for (M, K, N) in ((rows, dim, 32), (rows, dim, dim), (rows, nmid, dim), (Mp, nmid, nmid), (rows, dim, nmid),
                              (rows, self.cin, dim)):
                need = max(need, int(f(ops.C.c_longlong(M), K, N)))
            self._buf[key] = dict(
                rows=rows, R=R, Mb=Mb, Rb=Rb, acc_all=acc, units=units, x0=bf(R, dim), h=bf(R, dim), logits=bf(R, 32),
                dlogits=bf(R, 32), dh=bf(R, dim), dout=[bf(R, dim), bf(R, dim)], da3=bf(R, nmid), dc2b=bf(Rb, nmid),
                da2=bf(R, nmid), dc1=bf(R, nmid), da1=bf(R, dim),
                accb=torch.zeros((n, dim, 2), dtype=torch.float64, device=dev),
                ws=torch.empty(need, dtype=torch.uint8, device=dev),
                counts=torch.zeros((n, 2), dtype=torch.float32, device=dev))
        return self._buf[key]

    def forward(self, plane: types.FeaturePlane) -> Dict:
        f, valid = plane.features.contiguous(), plane.valid.contiguous()
        n, H, W, C = f.shape
        if C != self.cin:
            raise ValueError(f"the plane has {C} channels, layers_0 expects {self.cin}")
        buf = self._buffers(n, H, W)
        rows, dim, nmid, Bm = buf["rows"], self.dim, self.nmid, self.bank.b_mats
        self.bank.run()
        buf["acc_all"].zero_()
        ops.gemm(f.view(rows, C), Bm[self.d0], buf["x0"], m_rows=rows, bias=self.d0_b)            # nn.Dense (:154-158)
        ops.gn_stats(buf["x0"], n, H * W, dim, False, buf["units"][0]["acc"][0])
        hp, wp = H + 2, W + 2
        seg = [(a - 1) * wp + (b - 1) for a in range(3) for b in range(3)]
        x = buf["x0"]
        for i, (u, bu) in enumerate(zip(self.units, buf["units"])):                                 # resnet.py:103-134
            acc1, acc2, acc3 = bu["acc"]
            bu["x"] = x
            ops.gn_apply(x, n, H, W, dim, acc1, u["gn1"][0], u["gn1"][1], False, True, ops.LAYOUT_DENSE, bu["a1"])
            ops.gemm(bu["a1"], Bm[u["conv1"]], bu["y1"], m_rows=rows, gn_acc=acc2, gn_rows_per_img=H * W)
            ops.gn_apply(bu["y1"], n, H, W, nmid, acc2, u["gn2"][0], u["gn2"][1], False, True, ops.LAYOUT_PADDED, bu["a2"])
            ops.gemm(bu["a2"], Bm[u["conv2"]], bu["y2"], m_rows=buf["Mb"], seg_off=seg, seg_k=nmid,
                     remap=(hp, wp, 1, 1, H, W), gn_acc=acc3, gn_rows_per_img=H * W)
            ops.gn_apply(bu["y2"], n, H, W, nmid, acc3, u["gn3"][0], u["gn3"][1], False, True, ops.LAYOUT_DENSE, bu["a3"])
            nxt = buf["units"][i + 1]["acc"][0] if i + 1 < len(self.units) else None
            ops.gemm(bu["a3"], Bm[u["conv3"]], bu["out"], m_rows=rows, residual=x, gn_acc=nxt, gn_rows_per_img=H * W)
            x = bu["out"]
        buf["x_last"] = x
        ops.gemm(x, Bm[self.m0], buf["h"], m_rows=rows, bias=self.m0_b, relu=True)                  # MLP (:161)
        ops.gemm(buf["h"], Bm[self.m1], buf["logits"], m_rows=rows, bias=self.m1_b, row_mask=valid.view(rows))  # :186
        logits = buf["logits"][:rows, : self.num_classes].float().view(n, H, W, self.num_classes)
        pred = {"logits_areas": logits[..., : self.num_area]}
        if self.num_classes > self.num_area:
            rest = logits[..., self.num_area:]
            pred["logits_objects_exclusive"] = rest[..., : self.num_excl]
            pred["logits_objects_independent"] = rest[..., self.num_excl:]
        return pred

    # ------------------------------------------------------------------------------------------------------------
    def backward(self, plane: types.FeaturePlane, buf: Dict) -> None:
        """Gradients of every decoder parameter from buf['dlogits'] (bf16 [rows, 32]) into the `g` arrays."""
        f = plane.features.contiguous()
        n, H, W, C = f.shape
        rows, dim, nmid, Bm = buf["rows"], self.dim, self.nmid, self.bank.b_mats
        ws, g = buf["ws"], self.by_bank
        hp, wp = H + 2, W + 2
        seg = [(a - 1) * wp + (b - 1) for a in range(3) for b in range(3)]
        rec_of = {r["path"]: r for r in self.rec}
        # MLP: Dense_1, ReLU, Dense_0
        ops.dense_wgrad(buf["h"], buf["dlogits"], rows, dim, 32, g[self.m1]["g"], rec_of[("layers_3", "Dense_1", "bias")]["g"], ws)
        ops.cast_pad_bf16(g[self.m1]["p"], self.wc_m1)
        ops.gemm(buf["dlogits"], self.wc_m1, buf["dh"], m_rows=rows, seg_k=32)
        ops.relu_bwd(buf["h"], buf["dh"], rows * dim)
        ops.dense_wgrad(buf["x_last"], buf["dh"], rows, dim, dim, g[self.m0]["g"], rec_of[("layers_3", "Dense_0", "bias")]["g"], ws)
        ops.cast_pad_bf16(g[self.m0]["p"], self.wc_m0)
        cur = 0
        dout = buf["dout"][cur]
        ops.gemm(buf["dh"], self.wc_m0, dout, m_rows=rows, seg_k=dim)
        # residual units, last to first
        lo = wp + 1
        Mp = image_encoder._round_up(buf["Mb"] - 2 * lo, 16)
        for u, bu in zip(reversed(self.units), reversed(buf["units"])):
            acc1, acc2, acc3 = bu["acc"]
            r1, r2, r3 = g[u["conv1"]], g[u["conv2"]], g[u["conv3"]]
            # conv3 (1x1): dWs = a3^T dout, da3 = dout W3s^T
            ops.dense_wgrad(bu["a3"], dout, rows, nmid, dim, r3["gs"], None, ws)
            ops.wt_segments(Bm[u["conv3"]], dim, nmid, 1, u["bt3"])
            ops.gemm(dout, u["bt3"], buf["da3"], m_rows=rows, seg_k=dim)
            ops.gn_backward(bu["y2"], buf["da3"], n, H, W, nmid, acc3, u["gn3"][0], u["gn3"][1], buf["accb"], buf["dc2b"],
                            u["dgn3"][0], u["dgn3"][1], post_relu=True, padded_out=True)
            # conv2 (3x3): nine row-shifted products on the bordered layouts; 9-segment dX with mirrored offsets
            for t, off in enumerate(seg):
                ops.dense_wgrad(bu["a2"][lo + off: lo + off + Mp], buf["dc2b"][lo: lo + Mp], Mp, nmid, nmid,
                                r2["gs"][t * nmid: (t + 1) * nmid], None, ws)
            ops.wt_segments(Bm[u["conv2"]], nmid, nmid, 9, u["bt2"])
            ops.gemm(buf["dc2b"], u["bt2"], buf["da2"], m_rows=buf["Mb"], seg_off=[-o for o in seg], seg_k=nmid,
                     remap=(hp, wp, 1, 1, H, W))
            ops.gn_backward(bu["y1"], buf["da2"], n, H, W, nmid, acc2, u["gn2"][0], u["gn2"][1], buf["accb"], buf["dc1"],
                            u["dgn2"][0], u["dgn2"][1], post_relu=True)
            # conv1 (1x1)
            ops.dense_wgrad(bu["a1"], buf["dc1"], rows, dim, nmid, r1["gs"], None, ws)
            ops.wt_segments(Bm[u["conv1"]], nmid, dim, 1, u["bt1"])
            ops.gemm(buf["dc1"], u["bt1"], buf["da1"], m_rows=rows, seg_k=nmid)
            nxt = buf["dout"][1 - cur]
            ops.gn_backward(bu["x"], buf["da1"], n, H, W, dim, acc1, u["gn1"][0], u["gn1"][1], buf["accb"], nxt,
                            u["dgn1"][0], u["dgn1"][1], post_relu=True, add=dout)                    # + identity shortcut
            cur, dout = 1 - cur, nxt
            for r in (r1, r2, r3):
                ops.stdconv_backward(r["p"], r["gs"], r["g"])
        # layers_0
        ops.dense_wgrad(f.view(rows, C), dout, rows, C, dim, g[self.d0]["g"], rec_of[("layers_0", "bias")]["g"], ws)

    def train_step(self, plane: types.FeaturePlane, model, data: Dict, update: bool = True):
        """One training step on frozen BEV features.  Returns (per-example total loss, losses, metrics); gradients stay
        in the `g` arrays (`grads_tree`), averaged over ranks when torch.distributed is initialised."""
        from . import parallel
        pred = self.forward(plane)
        pred["bev_features"] = plane
        losses, metrics, ctx = model.loss_metrics_function(pred, data, return_context=True)
        n, H, W, _ = plane.features.shape
        buf = self._buffers(n, H, W)
        ops.sem_loss_grad(ctx["logits"], ctx["labels_area"], ctx["valid_area"], ctx["labels_excl"], ctx["masks_indep"],
                          ctx["valid"], self.num_area, self.num_excl, self.num_indep, ctx["weights"], buf["counts"],
                          buf["dlogits"])
        semantic_net.apply_batch_mask(buf["dlogits"], data, n)
        self.backward(plane, buf)
        self.bucket.allreduce_mean()                                             # jax.lax.pmean (trainer.py:231-234)
        if update:
            self.apply_update()
        return losses["total"], losses, metrics

    def apply_update(self) -> bool:
        """optax.adam on the fp32 masters, skipped when any (averaged) gradient is non-finite (`trainer.py:260-276`: the
        reference keeps the old parameters AND optimiser state for such a step).  Returns whether the step was applied."""
        if not bool(self.bucket.all_finite().item()):
            self.skipped_steps = getattr(self, "skipped_steps", 0) + 1
            return False
        self.step += 1
        for r, (m, v) in zip(self.rec, self.mom):
            ops.adam_step(r["p"].view(-1), m.view(-1), v.view(-1), r["g"].view(-1), self.lr, self.step)
        return True

    # ------------------------------------------------------------------------------------------------------------
    def _tree(self, key: str) -> Dict:
        out: Dict = {}
        for r in self.rec:
            t = r[key]
            if "bank" in r:
                a = t[:, : r["shape"][-1]].cpu().numpy().reshape(r["shape"]).copy()
            else:
                a = t[: r["n"]].cpu().numpy().reshape(r["shape"]).copy()
            d = out
            for k in r["path"][:-1]:
                d = d.setdefault(k, {})
            d[r["path"][-1]] = a
        return out

    def params_tree(self) -> Dict:
        """Current parameters as the Flax tree of the decoder (host, fp32)."""
        return self._tree("p")

    def grads_tree(self) -> Dict:
        return self._tree("g")
