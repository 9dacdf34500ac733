"""A NumPy/SciPy stand-in for the subset of the `jax` / `dataclass_array` / `etils` / `flax` APIs that the PURE
functions of /root/reference use, so that the reference's own source can be executed in an image without JAX.

Only used by tests/golden/make_golden.py (fixture generation, in the build container).  Semantics follow
SURVEY.md Appendix A; the stand-ins for jax.scipy.ndimage.map_coordinates / jax.scipy.signal.convolve /
lax.top_k / jax.nn.softmax are SciPy / NumPy (that is the part that stays "parity unpinned").
"""
from __future__ import annotations

import dataclasses
import sys
import types

import numpy as np
import scipy.ndimage
import scipy.signal

F = np.float32


# ------------------------------------------------------------------------------------------------
# jax.numpy: numpy with float32 defaults for freshly created floating arrays
# ------------------------------------------------------------------------------------------------
def _f32(a):
    a = np.asarray(a)
    return a.astype(F) if a.dtype == np.float64 else a


class _Jnp(types.ModuleType):
    ndarray = np.ndarray
    float32, bfloat16, int32, inf, pi, newaxis = np.float32, np.float32, np.int32, np.inf, np.pi, None

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def asarray(a, dtype=None):
        return _f32(np.asarray(a, dtype=dtype)) if dtype is None else np.asarray(a, dtype=dtype)

    array = asarray

    @staticmethod
    def linspace(*a, **k):
        return _f32(np.linspace(*a, **k))

    @staticmethod
    def arange(*a, **k):
        return _f32(np.arange(*a, **k))

    @staticmethod
    def zeros(shape, dtype=F):
        return np.zeros(shape, dtype=dtype).view(AtArray)

    @staticmethod
    def ones(shape, dtype=F):
        return np.ones(shape, dtype=dtype)

    @staticmethod
    def flip(a, axis=None):
        return np.flip(a, axis=axis)

    @staticmethod
    def _reduce(name):
        def f(a, *args, where=None, **kw):
            if isinstance(kw.get("axis"), list):  # jax accepts a list of axes
                kw["axis"] = tuple(kw["axis"])
            if where is not None:  # jax accepts any dtype as a mask
                kw["where"] = np.asarray(where).astype(bool)
            return getattr(np, name)(a, *args, **kw)
        return staticmethod(f)

    @staticmethod
    def where(c, a, b):
        r = np.where(c, a, b)
        scal = lambda v: isinstance(v, (int, float)) or (isinstance(v, np.ndarray) and v.ndim == 0 and v.dtype == np.float64)
        if r.dtype == np.float64 and (scal(a) or scal(b)):
            other = b if scal(a) else a
            if isinstance(other, np.ndarray) and other.dtype == F or (scal(a) and scal(b)):
                return r.astype(F)
        return r


for _n in ("mean", "var", "max", "min", "sum"):
    setattr(_Jnp, _n, _Jnp._reduce(_n))
jnp = _Jnp("jax.numpy")


class AtArray(np.ndarray):
    """ndarray with the functional update syntax of jax arrays: `x.at[idx].set(v)` returns an updated copy.  NumPy
    ufuncs propagate the subclass, so results of arithmetic / comparisons on AtArrays support `.at` too."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def set(self, v):
        out = np.array(self.a, copy=True).view(AtArray)
        out[self.idx] = v
        return out


class ClampArray(np.ndarray):
    """ndarray whose integer-array indexing clamps out-of-bounds indices, as JAX's NumPy-style indexing does
    ("out-of-bound indices are clamped" for retrieval)."""

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and all(isinstance(i, (int, np.integer, np.ndarray)) for i in idx):
            idx = tuple(np.clip(np.asarray(i), 0, n - 1) for i, n in zip(idx, self.shape))
        return np.asarray(np.ndarray.__getitem__(self, idx))


def clamp_indexing(a):
    return np.asarray(a).view(ClampArray)


# ------------------------------------------------------------------------------------------------
# dataclass_array: struct-of-arrays with a leading batch shape
# ------------------------------------------------------------------------------------------------
class _Spec:
    def __init__(self, spec):
        self.spec = spec

    @property
    def inner_rank(self):
        return len([t for t in self.spec.split() if t != "..."])

    def __or__(self, other):   # `FloatArray['N'] | tuple[...]` in return annotations
        return self

    __ror__ = __or__


class _ArrayType:
    def __getitem__(self, spec):
        return _Spec(spec)


class DataclassArray:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = {}
        for klass in reversed(cls.__mro__):
            for name, ann in getattr(klass, "__annotations__", {}).items():
                if isinstance(ann, _Spec):
                    fields[name] = ann
        cls._fields = fields

    def __init__(self, **kw):
        for name in self._fields:
            setattr(self, name, _f32(np.asarray(kw[name])).view(AtArray))

    @property
    def shape(self):
        name, spec = next(iter(self._fields.items()))
        a = getattr(self, name)
        return a.shape[: a.ndim - spec.inner_rank]

    @property
    def xnp(self):
        return jnp

    def __getitem__(self, idx):
        # the index addresses the BATCH dimensions only: expand an Ellipsis to the batch rank
        tup = idx if isinstance(idx, tuple) else (idx,)
        if any(i is Ellipsis for i in tup):
            k = sum(1 for i in tup if i is not Ellipsis and i is not None)
            pos = [j for j, i in enumerate(tup) if i is Ellipsis][0]
            tup = tup[:pos] + (slice(None),) * (len(self.shape) - k) + tup[pos + 1:]
        return type(self)(**{n: getattr(self, n)[tup] for n in self._fields})

    def __len__(self):
        return self.shape[0]

    def replace(self, **kw):
        d = {n: getattr(self, n) for n in self._fields}
        d.update(kw)
        return type(self)(**d)


# ------------------------------------------------------------------------------------------------
# jax.vmap / jit / checkpoint
# ------------------------------------------------------------------------------------------------
def _is_dca(x):
    return isinstance(x, DataclassArray)


def _axis_len(x, ax):
    if _is_dca(x):
        return x.shape[ax]
    return np.shape(x)[ax]


def _take(x, ax, i):
    if ax is None:
        return x
    if _is_dca(x):
        return type(x)(**{n: np.take(getattr(x, n), i, axis=ax if ax >= 0 else ax - x._fields[n].inner_rank)
                          for n in x._fields})
    if isinstance(x, (tuple, list)):
        return type(x)(_take(e, ax, i) for e in x)
    r = np.take(np.asarray(x), i, axis=ax)
    return r.view(ClampArray) if isinstance(x, ClampArray) else r


def _stack(items, ax):
    first = items[0]
    if isinstance(first, tuple):
        return tuple(_stack([it[k] for it in items], ax) for k in range(len(first)))
    if _is_dca(first):
        return type(first)(**{n: np.stack([getattr(it, n) for it in items], axis=ax if ax >= 0 else ax - first._fields[n].inner_rank)
                              for n in first._fields})
    return np.stack([np.asarray(it) for it in items], axis=ax)


def vmap(fn, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(_axis_len(a, ax) for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack(outs, out_axes)
    return wrapped


def jit(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


def checkpoint(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


# ------------------------------------------------------------------------------------------------
# jax.scipy / jax.lax / jax.nn pieces
# ------------------------------------------------------------------------------------------------
def map_coordinates(input, coordinates, order, mode="constant", cval=0.0):
    coords = [np.asarray(c, np.float64) for c in coordinates]
    scalar = coords[0].ndim == 0
    if scalar:
        coords = [c.reshape(1) for c in coords]
    out = scipy.ndimage.map_coordinates(np.asarray(input, dtype=np.float64), coords, order=order, mode=mode, cval=cval)
    if scalar:
        out = out.reshape(())
    return out.astype(np.asarray(input).dtype if np.asarray(input).dtype != bool else F)


def convolve(a, b, mode="full"):
    return scipy.signal.convolve(np.asarray(a, F), np.asarray(b, F), mode=mode, method="direct").astype(F)


def top_k(x, k):
    idx = np.argsort(-x, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(x, idx, -1), idx.astype(np.int32)


def softmax(x, axis=-1, where=None, initial=None):
    x = np.asarray(x)
    if where is None:
        m = x.max(axis=axis, keepdims=True)
        e = np.exp(x - m)
        return e / e.sum(axis=axis, keepdims=True)
    where = np.asarray(where).astype(bool)
    m = np.max(x, axis=axis, keepdims=True, where=where, initial=initial)
    e = np.where(where, np.exp(x - m), 0).astype(x.dtype)
    return (e / e.sum(axis=axis, keepdims=True)).astype(x.dtype)


def random_choice(rng, a, shape=(), replace=True, p=None):
    """Stand-in for jax.random.choice(key, a, shape, replace=True, p): inverse CDF `searchsorted(cumsum(p),
    total * (1 - u))` with `rng` a numpy Generator instead of a threefry key."""
    assert replace and p is not None
    cdf = np.cumsum(np.asarray(p, np.float64))
    u = rng.random(shape)
    return np.minimum(np.searchsorted(cdf, cdf[-1] * (1.0 - u), side="left"), a - 1)


def log_softmax(x, axis=-1):
    x = np.asarray(x)
    m = x.max(axis=axis, keepdims=True)
    return (x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True))).astype(x.dtype)


class _Permissive(types.ModuleType):
    """Module stub: any attribute is a permissive object (class / decorator / callable)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(name)


class _AnythingMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(name)


def _Anything(name):
    class _A(metaclass=_AnythingMeta):
        def __init__(self, *a, **k):
            pass

        def __new__(cls, *a, **k):
            if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], type):
                return a[0]  # used as a decorator
            return super().__new__(cls)

        def __call__(self, *a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return self

        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Anything(n)

        def __getitem__(self, k):
            return self

        def lock(self):
            return self
    _A.__name__ = name
    return _A


def install(reference_root: str, flax_modules: bool = False) -> None:
    """Register the stand-ins in sys.modules and the reference tree as bare packages (no __init__ execution)."""
    jax = types.ModuleType("jax")
    jax.numpy, jax.vmap, jax.jit, jax.checkpoint = jnp, vmap, jit, checkpoint
    jsp = types.ModuleType("jax.scipy")
    jnd = types.ModuleType("jax.scipy.ndimage"); jnd.map_coordinates = map_coordinates
    jsg = types.ModuleType("jax.scipy.signal"); jsg.convolve = convolve
    jsp.ndimage, jsp.signal = jnd, jsg
    lax = types.ModuleType("jax.lax"); lax.top_k = top_k
    jnn = _Permissive("jax.nn"); jnn.softmax = softmax
    jnn.initializers = _Permissive("jax.nn.initializers")
    jax.scipy, jax.lax, jax.nn = jsp, lax, jnn
    jax.random, jax.image, jax.tree_util = _Permissive("jax.random"), _Permissive("jax.image"), _Permissive("jax.tree_util")
    jax.random.choice = random_choice
    def _split(rng, n=2):   # an object array so that vmap indexes it like the key array jax.random.split returns
        keys = np.empty(n, dtype=object)
        for i in range(n):
            keys[i] = np.random.default_rng(int(rng.integers(1 << 62)))
        return keys
    jax.random.split = _split
    lax.stop_gradient = lambda x: x

    def tree_map(fn, *trees):
        first = trees[0]
        if _is_dca(first):
            return type(first)(**{n: fn(*[getattr(t, n) for t in trees]) for n in first._fields})
        if isinstance(first, dict):
            return {k: tree_map(fn, *[t[k] for t in trees]) for k in first}
        if isinstance(first, (list, tuple)):
            return type(first)(tree_map(fn, *parts) for parts in zip(*trees))
        return fn(*trees)
    jax.tree_util.tree_map = tree_map
    jnn.log_softmax = log_softmax
    jnn.relu = lambda x: np.maximum(x, 0)
    jnn.log_sigmoid = lambda x: (-(np.maximum(-np.asarray(x), 0) + np.log1p(np.exp(-np.abs(np.asarray(x)))))).astype(np.asarray(x).dtype)
    jnn.sigmoid = lambda x: (1 / (1 + np.exp(-np.asarray(x)))).astype(np.asarray(x).dtype)
    # optax (published definitions): logsumexp(logits) - logits[label]; -y log_sigmoid(x) - (1 - y) log_sigmoid(-x)
    optax = types.ModuleType("optax")

    def _sxent(logits, labels):
        l = np.asarray(logits)
        l = l - l.max(-1, keepdims=True)
        return (np.log(np.exp(l).sum(-1)) - np.take_along_axis(l, np.asarray(labels)[..., None], -1)[..., 0]).astype(l.dtype)

    def _bxent(logits, labels):
        x = np.asarray(logits)
        ls = lambda t: -(np.maximum(-t, 0) + np.log1p(np.exp(-np.abs(t))))
        y = np.asarray(labels).astype(x.dtype)
        return (-y * ls(x) - (1 - y) * ls(-x)).astype(x.dtype)
    optax.softmax_cross_entropy_with_integer_labels, optax.sigmoid_binary_cross_entropy = _sxent, _bxent
    mods = {"jax": jax, "jax.numpy": jnp, "jax.scipy": jsp, "jax.scipy.ndimage": jnd, "jax.scipy.signal": jsg,
            "jax.lax": lax, "jax.nn": jnn, "jax.nn.initializers": jnn.initializers, "jax.random": jax.random}
    dca = types.ModuleType("dataclass_array"); dca.DataclassArray = DataclassArray
    dca.utils = types.SimpleNamespace(np_utils=types.SimpleNamespace(get_xnp=lambda x: jnp))
    mods["dataclass_array"] = dca
    et = types.ModuleType("etils"); at = types.ModuleType("etils.array_types")
    at.BoolArray = at.FloatArray = at.IntArray = at.Array = _ArrayType()
    et.array_types = at; et.epath = _Permissive("etils.epath")
    mods.update({"etils": et, "etils.array_types": at, "etils.epath": et.epath})
    chex = types.ModuleType("chex"); chex.dataclass = dataclasses.dataclass
    mods["chex"] = chex
    for name in ("flax", "flax.linen", "flax.training", "flax.training.checkpoints", "ml_collections",
                 "ml_collections.config_dict", "tensorflow_datasets", "tensorflow", "scenic", "clu"):
        mods[name] = _Permissive(name)
    mods["optax"] = optax
    mods["flax.linen"].log_sigmoid = jnn.log_sigmoid
    mods["flax"].linen = mods["flax.linen"]
    mods["flax"].training = mods["flax.training"]
    mods["flax.training"].checkpoints = mods["flax.training.checkpoints"]
    mods["ml_collections"].config_dict = mods["ml_collections.config_dict"]
    if flax_modules:   # executable stand-ins for nn.Module / nn.Conv / ... (tests/golden/jaxshim/flaxshim.py)
        from . import flaxshim
        flaxshim.install(mods, jax)
    sys.modules.update(mods)
    import os
    for pkg in ("snap", "snap.models", "snap.utils", "snap.configs", "snap.data"):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(reference_root, *pkg.split("."))]
        sys.modules[pkg] = m
