"""Restatement of snap/models/layers.py. Test infrastructure."""
from __future__ import annotations

import numpy as np

F = np.float32


def normalize(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """snap/models/layers.py:45-52: L2-normalise the last axis; zero where ||x|| < eps."""
    x_ = x.astype(F)
    nrm = np.sqrt(np.sum(x_ * x_, axis=-1, keepdims=True)).astype(F)
    invalid = nrm < F(eps)
    y = np.where(invalid, F(eps), x_)
    nrm_y = np.sqrt(np.sum(y * y, axis=-1, keepdims=True)).astype(F)
    z = (x_ / nrm_y).astype(F)
    return np.where(invalid, F(0), z).astype(F)


def dense(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray | None) -> np.ndarray:
    """flax.linen.Dense: x @ kernel[in,out] + bias (fp32 accumulate)."""
    y = x.astype(F) @ kernel.astype(F)
    if bias is not None:
        y = y + bias.astype(F)
    return y.astype(F)


def mlp(x: np.ndarray, params: dict, apply_input_activation: bool = False, rd=lambda a: a) -> np.ndarray:
    """snap/models/layers.py:55-78 with activation='relu'.  params = {'Dense_0': {kernel,bias}, ...}.
    `rd` is the rounding hook applied to every layer output (identity = fp32, bf16 emulation otherwise)."""
    i = 0
    while f"Dense_{i}" in params:
        if i > 0 or apply_input_activation:
            x = np.maximum(x, 0)
        # flax.linen.Dense: y = dot_general(x, kernel) [materialised in dtype]; y += bias [again in dtype]
        x = rd(dense(x, params[f"Dense_{i}"]["kernel"], None))
        if params[f"Dense_{i}"].get("bias") is not None:
            x = rd(x + params[f"Dense_{i}"]["bias"].astype(F))
        i += 1
    return x
