"""Flax-msgpack checkpoint interchange (snap_b200/checkpoint.py): byte-level known answer built by hand from the
msgpack specification + the ext-type layout of flax/serialization.py, round trips, bf16 widening, chunked arrays, and
a full BEVLocalizer parameter tree through save / restore / check_tree."""
import struct

import os

import numpy as np
import pytest

from snap_b200 import checkpoint as ck

F = np.float32


def test_known_answer_bytes():
    arr = np.array([1.5, -2.0], F)
    # {'a': ndarray} -> fixmap(1), fixstr 'a', ext8(type 1) of msgpack((shape, dtype name, raw bytes))
    payload = b"\x93" + b"\x91\x02" + b"\xa7float32" + b"\xc4\x08" + struct.pack("<2f", 1.5, -2.0)
    expected = b"\x81" + b"\xa1a" + b"\xc7" + bytes([len(payload)]) + b"\x01" + payload
    assert ck.msgpack_serialize({"a": arr}) == expected
    back = ck.msgpack_restore(expected)
    assert back["a"].dtype == F and np.array_equal(back["a"], arr)
    # numpy scalar -> ext 3, python ints / floats stay native msgpack
    s = ck.msgpack_serialize({"step": np.int32(7), "lr": 0.5, "n": 3})
    assert b"\xc7" in s and ck.msgpack_restore(s) == {"step": 7, "lr": 0.5, "n": 3}
    assert ck.msgpack_restore(s)["step"].dtype == np.int32


def test_bfloat16_arrays_are_widened_exactly():
    vals = np.array([1.0, -0.3984375, 2.0 ** 100], F)     # representable in bf16
    raw = (vals.view(np.uint32) >> 16).astype(np.uint16)
    payload = ck.msgpack.packb(((3,), "bfloat16", raw.tobytes()), use_bin_type=True)
    enc = ck.msgpack.packb({"w": ck.msgpack.ExtType(1, payload)})
    out = ck.msgpack_restore(enc)["w"]
    assert out.dtype == F and np.array_equal(out, vals)


def test_chunked_arrays(monkeypatch):
    monkeypatch.setattr(ck, "MAX_CHUNK_SIZE", 64)
    a = np.arange(100, dtype=F).reshape(4, 25)
    enc = ck.msgpack_serialize({"big": a, "small": np.ones(3, F)})
    raw = ck.msgpack.unpackb(enc, ext_hook=ck._ext_unpack, raw=False)
    assert raw["big"][ck.CHUNK_KEY] is True and len(raw["big"]["chunks"]) == 7 and raw["big"]["shape"] == {"0": 4, "1": 25}
    out = ck.msgpack_restore(enc)
    assert np.array_equal(out["big"], a) and np.array_equal(out["small"], np.ones(3, F))


def test_train_state_round_trip_and_tree_check(tmp_path):
    from snap_b200 import bev_localizer, configs, params, types
    rng = np.random.default_rng(0)
    cfg = configs.bev_localizer()
    cfg.bev_mapper = configs.bev_mapper(("streetview", "aerial"))
    cfg.filter_points_in_fov = True
    loc = bev_localizer.BEVLocalizer(cfg, None, types.Grid2D((64, 64), 0.2))
    tree = loc.init_params(params.init_bev_mapper(rng, cfg.bev_mapper))
    state = {"global_step": np.int32(1234), "params": tree, "model_state": {}, "rng": np.array([0, 1], np.uint32),
             "opt_state": {"0": {"count": np.int32(1234)}}, "metadata": {"lr": 1e-3}}
    for step in (9, 10, 1234):
        path = ck.save_checkpoint(str(tmp_path), {**state, "global_step": np.int32(step)}, step, keep=3)
    assert ck.latest_checkpoint(str(tmp_path)) == path and path.endswith("checkpoint_1234")   # natural sort: 1234 > 10 > 9
    # flax's legacy rules: an equal or earlier step is refused unless overwrite (which drops the later checkpoints) ...
    with pytest.raises(ck.InvalidCheckpointError):
        ck.save_checkpoint(str(tmp_path), state, 1234, keep=3)
    with pytest.raises(ck.InvalidCheckpointError):
        ck.save_checkpoint(str(tmp_path), state, 100, keep=3)
    # ... and only the `keep` newest checkpoints survive a save
    d2 = tmp_path / "rolling"
    for step in (1, 2, 3, 4):
        ck.save_checkpoint(str(d2), {"global_step": np.int32(step)}, step, keep=2)
    assert sorted(os.listdir(d2)) == ["checkpoint_3", "checkpoint_4"]
    ck.save_checkpoint(str(d2), {"global_step": np.int32(3)}, 3, keep=2, overwrite=True)
    assert sorted(os.listdir(d2)) == ["checkpoint_3"] and ck.restore_checkpoint(str(d2))["global_step"] == 3
    back = ck.restore_checkpoint(str(tmp_path))
    assert back["global_step"] == 1234 and back["metadata"] == {"lr": 1e-3} and back["model_state"] == {}
    p = ck.load_params(str(tmp_path))
    ck.check_tree(p, tree)
    # Appendix B names and layouts
    enc = p["bev_mapper"]["streetview_encoder"]["image_encoder"]["encoder"]
    assert enc["root_block"]["conv_root"]["kernel"].shape == (7, 7, 3, 64)
    assert enc["block1"]["unit01"]["conv_proj"]["kernel"].shape == (1, 1, 64, 256)
    assert enc["block4"]["unit03"]["gn3"]["scale"].shape == (1, 1, 1, 512)
    assert p["bev_mapper"]["streetview_encoder"]["fusion_mlp"]["Dense_0"]["kernel"].shape == (257, 256)
    assert p["bev_mapper"]["aerial_encoder"]["encoder"]["conv_root"]["kernel"].shape == (3, 3, 3, 64)
    assert p["bev_mapper"]["matching_proj"]["kernel"].shape == (128, 32) and p["temperature"].shape == ()
    flat_a, flat_b = [], []

    def walk(t, out):
        for k in sorted(t):
            walk(t[k], out) if isinstance(t[k], dict) else out.append(np.asarray(t[k]))
    walk(tree, flat_a); walk(p, flat_b)
    assert len(flat_a) == len(flat_b) > 300 and all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(flat_a, flat_b))
    bad = {**p, "bev_mapper": {k: v for k, v in p["bev_mapper"].items() if k != "matching_proj"}}
    with pytest.raises(KeyError):
        ck.check_tree(bad, tree)
    assert ck.load_params(str(tmp_path), step=9)["temperature"] == F(2.0)
    older = ck.save_checkpoint(str(tmp_path / "old"), {"optimizer": {"target": {"params": {"temperature": F(1.0)}}}}, 1)
    assert ck.load_params(older)["temperature"] == F(1.0)
