"""Checkpoint interchange with the reference: the legacy Flax msgpack format that
`flax.training.checkpoints.save_checkpoint` writes for `snap.train` (`snap/train.py:22,34` disables orbax;
`snap/trainer.py:437-440,595-603` restore / save through scenic's `train_utils`).

A checkpoint file `<workdir>/checkpoint_<step>` is `flax.serialization.msgpack_serialize(to_state_dict(train_state))`:
a msgpack map whose leaves are msgpack *extension* objects
    ext 1 (ndarray)  = msgpack((shape, dtype.name, C-order bytes), use_bin_type=True)
    ext 3 (npscalar) = the same for a 0-d array;  ext 2 (native complex) = msgpack((real, imag))
arrays above 2**30 bytes are split into {'__msgpack_chunked_array__': True, 'shape': {'0': ..}, 'chunks': {'0': ..}}.
scenic's TrainState serialises as {'global_step', 'opt_state', 'params', 'model_state', 'rng', 'metadata'}; the
parameter tree below 'params' carries exactly the names of SURVEY Appendix B, which `snap_b200.params` also uses, so a
restored tree plugs into `BEVMapper` / `BEVLocalizer` / `SemanticNet` unchanged.

flax is not installed in this image: the format is restated from its published source (flax/serialization.py) and is
*parity unpinned* beyond the byte-level known-answer test in tests/test_checkpoint.py.  Host-side only (no GPU work).
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, Optional

import msgpack
import numpy as np

EXT_NDARRAY, EXT_NATIVE_COMPLEX, EXT_NPSCALAR = 1, 2, 3
MAX_CHUNK_SIZE = 2 ** 30
CHUNK_KEY = "__msgpack_chunked_array__"


def _ndarray_to_bytes(arr: np.ndarray, dtype_name: Optional[str] = None) -> bytes:
    if arr.dtype.hasobject or arr.dtype.isalignedstruct:
        raise ValueError("Object and structured dtypes not supported for serialization of ndarrays.")
    return msgpack.packb((arr.shape, dtype_name or arr.dtype.name, arr.tobytes("C")), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
    shape, dtype_name, buffer = msgpack.unpackb(data, raw=True)
    name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
    if name == "bfloat16":   # NumPy has no bfloat16: widen to float32 (exact)
        raw = np.frombuffer(buffer, dtype=np.uint16).astype(np.uint32) << 16
        return raw.view(np.float32).reshape(shape)
    return np.frombuffer(buffer, dtype=np.dtype(name)).reshape(shape, order="C")


def _ext_pack(x: Any):
    if isinstance(x, np.ndarray):
        return msgpack.ExtType(EXT_NDARRAY, _ndarray_to_bytes(x))
    if isinstance(x, np.generic):
        return msgpack.ExtType(EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
    if isinstance(x, complex):
        return msgpack.ExtType(EXT_NATIVE_COMPLEX, msgpack.packb((x.real, x.imag)))
    return x


def _ext_unpack(code: int, data: bytes):
    if code == EXT_NDARRAY:
        return _ndarray_from_bytes(data)
    if code == EXT_NATIVE_COMPLEX:
        re_, im = msgpack.unpackb(data)
        return complex(re_, im)
    if code == EXT_NPSCALAR:
        return _ndarray_from_bytes(data)[()]
    return msgpack.ExtType(code, data)


def _tuple_to_dict(tpl):
    return {str(i): v for i, v in enumerate(tpl)}


def _chunk(arr: np.ndarray) -> Dict:
    chunksize = max(1, int(MAX_CHUNK_SIZE / arr.dtype.itemsize))
    flat = arr.reshape(-1)
    chunks = [flat[i:i + chunksize] for i in range(0, flat.size, chunksize)]
    return {CHUNK_KEY: True, "shape": _tuple_to_dict(arr.shape), "chunks": _tuple_to_dict(chunks)}


def _prepare(tree: Any) -> Any:
    """numpy-convert leaves, chunk the huge ones, tuples/lists -> state-dict maps with string keys."""
    if isinstance(tree, dict):
        return {str(k): _prepare(v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return {str(i): _prepare(v) for i, v in enumerate(tree)}
    if hasattr(tree, "detach") and hasattr(tree, "cpu"):   # torch tensor
        t = tree.detach().cpu()
        tree = t.float().numpy() if str(t.dtype) == "torch.bfloat16" else t.numpy()
    if isinstance(tree, np.ndarray) and tree.size * tree.dtype.itemsize > MAX_CHUNK_SIZE:
        return _chunk(tree)
    return tree


def _unchunk(tree: Any) -> Any:
    if isinstance(tree, dict):
        if CHUNK_KEY in tree:
            shape = tuple(tree["shape"][str(i)] for i in range(len(tree["shape"])))
            chunks = [tree["chunks"][str(i)] for i in range(len(tree["chunks"]))]
            return np.concatenate(chunks).reshape(shape)
        return {k: _unchunk(v) for k, v in tree.items()}
    return tree


def msgpack_serialize(pytree: Any) -> bytes:
    """`flax.serialization.msgpack_serialize` of a state dict (nested dicts of arrays / scalars)."""
    return msgpack.packb(_prepare(pytree), default=_ext_pack, strict_types=True)


def msgpack_restore(encoded: bytes) -> Any:
    """`flax.serialization.msgpack_restore`."""
    return _unchunk(msgpack.unpackb(encoded, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ------------------------------------------------------------------------------------------------------------------
# flax.training.checkpoints (legacy, non-orbax): files `<prefix><step>` in a directory
# ------------------------------------------------------------------------------------------------------------------
def _natural_key(name: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def latest_checkpoint(ckpt_dir: str, prefix: str = "checkpoint_") -> Optional[str]:
    if not os.path.isdir(ckpt_dir):
        return None
    names = [n for n in os.listdir(ckpt_dir) if n.startswith(prefix) and not n.endswith("tmp") and n[len(prefix):].isdigit()]
    return os.path.join(ckpt_dir, sorted(names, key=_natural_key)[-1]) if names else None


class InvalidCheckpointError(Exception):
    """`flax.errors.InvalidCheckpointError`: a checkpoint at an equal or later step already exists."""


def _steps(ckpt_dir: str, prefix: str):
    if not os.path.isdir(ckpt_dir):
        return []
    return sorted(int(n[len(prefix):]) for n in os.listdir(ckpt_dir) if n.startswith(prefix) and n[len(prefix):].isdigit())


def save_checkpoint(ckpt_dir: str, state: Dict, step: int, prefix: str = "checkpoint_", keep: int = 1,
                    overwrite: bool = False) -> str:
    """Write `state` (e.g. {'params': ..., 'global_step': ...}) as `<ckpt_dir>/<prefix><step>` (atomic rename), with the
    rules of the legacy `flax.training.checkpoints.save_checkpoint` (restated from flax's published source; flax is not
    installed here): saving at a step that is not later than the newest existing checkpoint raises unless `overwrite`
    (which then removes the checkpoints at later steps), and only the `keep` newest checkpoints are retained.
    bf16 tensors are written widened to float32 (exact); a loaded bf16 checkpoint therefore comes back as float32."""
    os.makedirs(ckpt_dir, exist_ok=True)
    path = os.path.join(ckpt_dir, f"{prefix}{step}")
    steps = _steps(ckpt_dir, prefix)
    if steps and step <= steps[-1]:
        if not overwrite:
            raise InvalidCheckpointError(f"trying to save an outdated checkpoint at step {step} <= latest step {steps[-1]} "
                                         f"in {ckpt_dir} (pass overwrite=True to replace it and drop the later ones)")
        for st in steps:
            if st > step:
                os.remove(os.path.join(ckpt_dir, f"{prefix}{st}"))
    tmp = path + "tmp"
    with open(tmp, "wb") as f:
        f.write(msgpack_serialize(state))
    os.replace(tmp, path)
    if keep is not None and keep > 0:
        for st in _steps(ckpt_dir, prefix)[:-keep]:
            os.remove(os.path.join(ckpt_dir, f"{prefix}{st}"))
    return path


def restore_checkpoint(ckpt_dir_or_file: str, step: Optional[int] = None, prefix: str = "checkpoint_") -> Optional[Dict]:
    path = ckpt_dir_or_file
    if os.path.isdir(path):
        path = os.path.join(path, f"{prefix}{step}") if step is not None else latest_checkpoint(path, prefix)
    if path is None or not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        return msgpack_restore(f.read())


def load_params(ckpt_dir_or_file: str, step: Optional[int] = None) -> Dict:
    """The parameter tree of a `snap.train` checkpoint: scenic's TrainState keeps it under 'params' (newer) or
    under 'optimizer'/'target' (older flax.optim states, scenic `pretrain_utils`)."""
    state = restore_checkpoint(ckpt_dir_or_file, step)
    if state is None:
        raise FileNotFoundError(f"no checkpoint under {ckpt_dir_or_file}")
    if "params" in state and state["params"] is not None:
        return state["params"]
    if "optimizer" in state and "target" in state["optimizer"]:
        tgt = state["optimizer"]["target"]
        return tgt.get("params", tgt)
    return state


def check_tree(params: Dict, like: Dict, path: str = "") -> None:
    """Raise if `params` does not have the names and shapes of `like` (e.g. `params.init_bev_mapper(...)`)."""
    missing = sorted(set(like) - set(params))
    extra = sorted(set(params) - set(like))
    if missing or extra:
        raise KeyError(f"parameter tree mismatch at '{path}': missing {missing}, unexpected {extra}")
    for k, v in like.items():
        if isinstance(v, dict):
            if not isinstance(params[k], dict):
                raise KeyError(f"'{path}/{k}' should be a sub-tree")
            check_tree(params[k], v, f"{path}/{k}")
        elif tuple(np.shape(params[k])) != tuple(np.shape(v)):
            raise ValueError(f"'{path}/{k}': shape {np.shape(params[k])}, expected {np.shape(v)}")
