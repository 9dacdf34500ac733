"""Semantic head forward vs oracle/semantic_net.py (bf16-emulation mode)."""
import numpy as np
import pytest
import torch

from util import F, bf16_np, rd_bf16, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_semantic_head_vs_oracle():
    from oracle import semantic_net as osn
    from snap_b200 import configs, params, semantic_net, types
    rng = np.random.default_rng(5)
    cfg = configs.semantic_net()
    p = params.round_to_bf16(params.perturb_affine(rng, params.init_semantic_decoder(rng, cfg)))
    B, G = 2, 24
    feats = bf16_np(rng.standard_normal((B, G, G, 128)))
    valid = rng.random((B, G, G)) > 0.2
    feats *= valid[..., None]
    ref = osn.semantic_decoder(feats, valid, p, rd_bf16)
    ref32 = osn.semantic_decoder(feats, valid, p)
    head = semantic_net.SemanticHead(cfg)
    plane = types.FeaturePlane(torch.from_numpy(feats).to(torch.bfloat16).cuda(), torch.from_numpy(valid.astype(np.uint8)).cuda())
    pred = head.apply({"params": {"decoder": p}}, plane)
    torch.cuda.synchronize()
    got = torch.cat([pred["logits_areas"], pred["logits_objects_exclusive"], pred["logits_objects_independent"]], -1).cpu().numpy()
    assert got.shape == ref.shape == (B, G, G, 12) and got.dtype == np.float32
    assert pred["logits_areas"].shape[-1] == 5 and pred["logits_objects_exclusive"].shape[-1] == 4
    assert not got[~valid].any()
    e, e32, eref = rel_l2(got, ref), rel_l2(got, ref32), rel_l2(ref, ref32)
    print(f"semantic head: vs bf16-oracle {e:.4f}, vs fp32-oracle {e32:.4f}, bf16-oracle vs fp32-oracle {eref:.4f}")
    # two residual units + 3 dense layers with identical rounding points: bf16 flips only
    assert e < 1e-2
