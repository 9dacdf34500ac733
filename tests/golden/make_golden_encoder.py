"""Golden fixtures for the ENCODER modules: the reference's OWN `resnet.ResNetV2`, `resnet.ResNetStage`,
`image_encoder.FPNDecoder`, `image_encoder.ImageEncoder` and `layers.MLP` (snap/models/resnet.py:46-216,
image_encoder.py:40-144, layers.py:55-78) executed on NumPy / torch-CPU under the stand-ins for jax and flax.linen
(tests/golden/jaxshim/flaxshim.py) with a parameter tree of the Flax layout.  Run in the build
container only:

    python tests/golden/make_golden_encoder.py     # writes tests/golden/encoder_modules.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")

from snap_b200 import configs, params  # noqa: E402  (parameter initialisers with the Flax tree names)

if __name__ == "__main__":
    import jaxshim  # noqa: E402
    jaxshim.install(REF, flax_modules=True)
    from snap.models import image_encoder as ie, layers, resnet  # noqa: E402

F = np.float32
rng = np.random.default_rng(2024)


class Cfg(dict):
    __getattr__ = dict.__getitem__


def flat(tree, pre=""):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from flat(v, pre + k + "/")
        else:
            yield pre + k, np.asarray(v)


out = {}


def encoder_case(tag):
    """Parameters are NOT stored: tests/test_golden.py re-creates them from the same seeds (snap_b200.params initialisers)."""
    skip_root, hw, seed = {"sv": (False, (40, 56), 501), "aerial": (True, (12, 16), 502)}[tag]
    enc_cfg = configs.image_encoder()
    enc_cfg.encoder.depth = [1, 1, 2, 1]
    enc_cfg.encoder.width = 0.5
    enc_cfg.encoder.skip_root_block = skip_root
    prng = np.random.default_rng(seed)
    p = params.perturb_affine(prng, params.init_image_encoder(prng, enc_cfg))
    img = np.random.default_rng(seed + 50).random((2, *hw, 3)).astype(F)
    return enc_cfg, p, img, skip_root


def stage_case():
    prng = np.random.default_rng(503)
    sp = {}
    for u in range(2):
        sp[f"unit{u + 1:02d}"] = {g: {"scale": (1 + 0.2 * prng.standard_normal((1, 1, 1, c))).astype(F), "bias": (0.2 * prng.standard_normal((1, 1, 1, c))).astype(F)}
                                  for g, c in (("gn1", 128), ("gn2", 32), ("gn3", 32))}
        sp[f"unit{u + 1:02d}"].update(conv1={"kernel": params.lecun_normal(prng, (1, 1, 128, 32))}, conv2={"kernel": params.lecun_normal(prng, (3, 3, 32, 32))},
                                      conv3={"kernel": params.lecun_normal(prng, (1, 1, 32, 128))})
    return sp, prng.standard_normal((2, 6, 5, 128)).astype(F)


def mlp_case():
    prng = np.random.default_rng(504)
    return params.perturb_affine(prng, params.init_mlp(prng, 10, (7, 4))), prng.standard_normal((5, 10)).astype(F)


def semantic_case(decoder_type):
    prng = np.random.default_rng(505 if decoder_type == "mlp" else 506)
    cfg = configs.semantic_net()
    cfg.decoder_type = decoder_type
    cfg.decoder_dim, cfg.resnet_num_units, cfg.mlp_num_layers = 128, 2, 2
    if decoder_type == "mlp":
        p = params.perturb_affine(prng, params.init_mlp(prng, 128, (128, 128, 12)))
    else:
        p = params.perturb_affine(prng, params.init_semantic_decoder(prng, cfg))
    feats = prng.standard_normal((2, 6, 5, 128)).astype(F)
    valid = prng.random((2, 6, 5)) < 0.8
    return cfg, p, feats * valid[..., None], valid


if __name__ == "__main__":
    import types as pytypes
    import jaxshim.flaxshim as fs
    from snap.models import semantic_net as sn, types as rtypes
    for dt in ("mlp", "resnet_stage"):       # SemanticNet.__call__ (semantic_net.py:167-198) with a stand-in bev_mapper
        scfg, p, feats, valid = semantic_case(dt)
        rcfg = Cfg(bev_mapper=Cfg(pretrained_path=None), decoder_type=dt, decoder_dim=128, mlp_num_layers=2, resnet_num_units=2,
                   apply_random_flip=False, area_classes=scfg.area_classes, object_classes_exclusive=scfg.object_classes_exclusive,
                   object_classes_independent=scfg.object_classes_independent)
        m = sn.SemanticNet(rcfg, None, F)
        m._params, m._counters = {"decoder": p}, {}
        fs._STACK.append(m)
        m.setup()
        fs._STACK.pop()
        m.decoder._parent = m
        object.__setattr__(m.decoder, "name", "decoder")
        m.bev_mapper = lambda data, train: {"bev_features": rtypes.FeaturePlane(features=feats, valid=valid)}
        pred = sn.SemanticNet.__call__(m, {"map": {}}, False)
        out[f"sem_{dt}_areas"], out[f"sem_{dt}_excl"] = pred["logits_areas"], pred["logits_objects_exclusive"]
        out[f"sem_{dt}_indep"] = pred["logits_objects_independent"]
        print("semantic", dt, np.asarray(pred["logits_areas"]).shape, np.asarray(pred["logits_objects_exclusive"]).shape)
    for tag in ("sv", "aerial"):
        enc_cfg, p, img, skip_root = encoder_case(tag)
        cfg = Cfg(encoder_name="resnet", output_dim=128, num_pyr_levels=None,
                  encoder=Cfg(width=0.5, depth=[1, 1, 2, 1], limit_num_blocks=4, skip_root_block=skip_root, checkpoint_blocks=False,
                              checkpoint_units=False, pretrained_path=None))
        m = ie.ImageEncoder(cfg, F)
        m._params, m._counters = p, {}
        fs._STACK.append(m)          # setup()-style module: bind the submodules by attribute name like flax does
        m.setup()
        fs._STACK.pop()
        m.encoder._parent, m.decoder._parent = m, m
        object.__setattr__(m.encoder, "name", "encoder")
        object.__setattr__(m.decoder, "name", "decoder")
        pyr = ie.ImageEncoder.__call__(m, img, False)
        for k, f in enumerate(pyr.features):
            out[f"{tag}_feat{k}"] = f
        print(tag, [np.asarray(f).shape for f in pyr.features], float(np.abs(pyr.features[-1]).mean()))
    sp, x = stage_case()
    y, units = resnet.ResNetStage(2, dtype=F).apply({"params": sp}, x)
    out["stage_y"], out["stage_unit01"] = y, units["unit01"]
    mp, xm = mlp_case()
    for act in (False, True):
        out[f"mlp_y{int(act)}"] = layers.MLP(Cfg(activation="relu", layers=(7, 4), apply_input_activation=act), F).apply({"params": mp}, xm)
    np.savez_compressed(os.path.join(HERE, "encoder_modules.npz"), **{k: np.asarray(v) for k, v in out.items()})
    print("saved", len(out), "arrays")
