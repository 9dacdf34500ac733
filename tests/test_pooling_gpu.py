"""VerticalPooling modes other than the fused 'max' (bev_mapper.py:56-88), modality fusion through them and the
bev_confidence head (bev_mapper.py:292-295) vs the oracle on identical bf16 inputs."""
import numpy as np
import pytest
import torch

from util import F, bf16_np, rd_bf16, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=F))


def _volume(rng, cells, Z, C=128):
    f = bf16_np(rng.standard_normal((cells, Z, C)))
    v = rng.random((cells, Z)) > 0.6
    v[: cells // 8] = False        # fully invalid columns (double-where branch)
    v[cells // 8: cells // 4] = True
    return f, v


@pytest.mark.parametrize("mode", ["max", "sum", "mean", "softmax", "weighted", "mlp"])
@pytest.mark.parametrize("Z", [60, 2])
def test_vertical_pooling_modes_vs_oracle(mode, Z):
    """Z=60: the scene column; Z=2: the modality axis of fuse_neural_maps (:247-252)."""
    from oracle import bev_mapper as obm
    from snap_b200 import bev_mapper, configs, params, types
    rng = np.random.default_rng(31 + Z)
    cells, C = 1000, 128
    f, v = _volume(rng, cells, Z)
    cfg = configs.vertical_pooling()
    cfg.pooling = mode
    p = {}
    if mode in ("softmax", "weighted"):
        p = {"confidence_head": {"kernel": bf16_np(rng.standard_normal((C, 1)) * 0.3), "bias": bf16_np(np.array([0.25]))}}
    elif mode == "mlp":
        p = {"fusion_mlp": params.round_to_bf16(params.perturb_affine(rng, params.init_mlp(rng, Z * C, (256, 128))))}
    dev = "cuda"
    vol = types.FeatureVolume(features=_t(f).to(torch.bfloat16).to(dev).view(10, 100, Z, C),
                              valid=torch.from_numpy(v.astype(np.uint8)).to(dev).view(10, 100, Z))
    pred = bev_mapper.VerticalPooling(cfg).apply({"params": p}, vol)
    torch.cuda.synchronize()
    ref = obm.vertical_pooling(f, v, mode, p, rd=obm.np_rd(rd_bf16))
    rp, rv = ref["plane"]
    got = pred["plane"].features.float().cpu().numpy().reshape(cells, C)
    assert np.array_equal(pred["plane"].valid.cpu().numpy().reshape(-1).astype(bool), rv), "valid plane differs"
    assert not got[~rv].any(), "columns without a valid level must be zero (:86)"
    if mode == "max":
        assert np.array_equal(got, rp)
    else:
        e = rel_l2(got[rv], rp[rv])
        print(f"{mode} Z={Z}: rel_l2 {e:.2e}")
        # one bf16 rounding of an fp32 accumulation on identical inputs: the summation order flips a few roundings
        assert e < 1e-3            # measured 0 (sum, mean) .. 2e-5 (softmax, weighted) .. 3.8e-4 (mlp)
    if mode in ("softmax", "weighted"):
        s, w = pred["scores"].cpu().numpy().reshape(cells, Z), pred["weights"].cpu().numpy().reshape(cells, Z)
        # logits are bf16 values of an fp32 dot: a flipped rounding moves one by a bf16 ulp (2^-8 relative)
        assert np.abs(s - ref["scores"]).max() <= 2.0 ** -7 * np.abs(ref["scores"]).max() + 1e-3
        assert np.abs(w - ref["weights"]).max() < 2e-2 and not w[~v].any()
        assert np.allclose(w[rv].sum(-1), 1.0, atol=1e-5)


def test_bev_confidence_and_non_max_modality_fusion_vs_oracle():
    from oracle import bev_mapper as obm
    from snap_b200 import bev_mapper, configs, ops, types
    rng = np.random.default_rng(41)
    cells, C = 4096, 128
    dev = "cuda"
    a, b = bf16_np(rng.standard_normal((cells, C))), bf16_np(rng.standard_normal((cells, C)))
    va, vb = rng.random(cells) > 0.4, rng.random(cells) > 0.2
    head = {"layers_0": {"kernel": bf16_np(rng.standard_normal((C, 1)) * 0.3), "bias": bf16_np(np.array([-0.5]))}}
    out = torch.empty(cells, dtype=torch.float32, device=dev)
    ops.confidence(_t(a).to(torch.bfloat16).to(dev), torch.from_numpy(va.astype(np.uint8)).to(dev), cells, C,
                   _t(head["layers_0"]["kernel"].reshape(-1)).to(dev), float(head["layers_0"]["bias"][0]), out)
    torch.cuda.synchronize()
    ref = obm.bev_confidence(a, va, head, rd=obm.np_rd(rd_bf16))
    got = out.cpu().numpy()
    assert not got[~va].any()
    assert np.abs(got - ref).max() <= 2.0 ** -7 * np.abs(ref).max() + 1e-3   # one bf16 ulp of the logit
    # modality fusion with 'mean' instead of the default 'max'
    cfg = configs.bev_mapper(("streetview", "aerial"))
    cfg.modality_fusion.pooling = "mean"
    mapper = bev_mapper.BEVMapper(cfg, types.Grid2D((64, 64), 0.2))
    mk = lambda f, v: types.FeaturePlane(features=_t(f).to(torch.bfloat16).to(dev).view(1, 64, 64, C),
                                         valid=torch.from_numpy(v.astype(np.uint8)).to(dev).view(1, 64, 64))
    fused = mapper.fuse_neural_maps([mk(a, va), mk(b, vb)], params={})
    torch.cuda.synchronize()
    rp, rv = obm.vertical_pooling(np.stack([a, b], -2), np.stack([va, vb], -1), "mean", rd=obm.np_rd(rd_bf16))["plane"]
    assert np.array_equal(fused.valid.cpu().numpy().reshape(-1).astype(bool), rv)
    assert rel_l2(fused.features.float().cpu().numpy().reshape(cells, C), rp) < 3e-3
