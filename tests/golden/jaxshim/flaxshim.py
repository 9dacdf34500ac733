"""A minimal stand-in for the subset of `flax.linen` that the reference's encoder modules use (nn.Module with dataclass
fields, @nn.compact, self.param, nn.Conv, nn.Dense, nn.max_pool, nn.relu, nn.remat, nn.Sequential, automatic submodule
names), so that `snap/models/resnet.py`, `image_encoder.FPNDecoder`, `layers.MLP` can be EXECUTED on NumPy/torch-CPU with
a given parameter tree.  Only used by tests/golden/make_golden_encoder.py.  The semantics of the flax primitives themselves
(Conv padding 'SAME', Dense, max_pool with -inf padding, auto-naming ClassName_i) follow SURVEY Appendix A and stay
'parity unpinned'; what gets pinned is the reference-authored module logic built on top of them."""
import types

import numpy as np
import torch
import torch.nn.functional as Fnn

F = np.float32
_STACK = []          # modules whose __call__ is running (innermost last)
_UNSET = object()


def _fields_of(cls):
    fields = []
    for klass in reversed(cls.__mro__):
        for name in getattr(klass, "__annotations__", {}):
            if name not in fields and not name.startswith("_"):
                fields.append(name)
    return fields


class Module:
    name: str = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        call = cls.__dict__.get("__call__")
        if call is not None and not getattr(call, "_wrapped", False):
            def wrapped(self, *a, __call=call, **k):
                self._begin()
                try:
                    return __call(self, *a, **k)
                finally:
                    _STACK.pop()
            wrapped._wrapped = True
            cls.__call__ = wrapped

    def __init__(self, *args, **kw):
        fields = [f for f in _fields_of(type(self)) if f != "name"]
        vals = dict(zip(fields, args))
        vals.update(kw)
        for f in fields + ["name"]:
            if f in vals:
                object.__setattr__(self, f, vals[f])
            elif not hasattr(type(self), f):
                raise TypeError(f"{type(self).__name__}: missing field {f}")
        unknown = set(vals) - set(fields) - {"name"}
        if unknown:
            raise TypeError(f"{type(self).__name__}: unknown fields {unknown}")
        self._parent = _STACK[-1] if _STACK else None
        self._params = None
        if self._parent is not None and self.name is None:       # flax auto-naming: ClassName_<index per class>
            k = self._parent._counters.get(type(self).__name__, 0)
            self._parent._counters[type(self).__name__] = k + 1
            object.__setattr__(self, "name", f"{type(self).__name__}_{k}")
        self.__post_init__()

    def __post_init__(self):
        pass

    def _begin(self):
        if self._params is None:
            if self._parent is None:
                raise RuntimeError("unbound module: use .apply(variables, ...)")
            self._params = self._parent._params[self.name]
        self._counters = {}
        _STACK.append(self)

    def apply(self, variables, *a, **k):
        self._params = variables["params"] if "params" in variables else variables
        return self(*a, **k)

    def param(self, name, init=None, *shape, **kw):
        return np.asarray(self._params[name], F)


def compact(fn):
    return fn


def remat(x, **kw):
    return x


def relu(x):
    return np.maximum(x, 0)


def _same_pad(n, k, s):
    total = max((-(-n // s) - 1) * s + k - n, 0)
    return total // 2, total - total // 2


class Conv(Module):
    features: int
    kernel_size: tuple
    strides: tuple = (1, 1)
    padding: object = "SAME"
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None
    param_dtype: object = None
    dtype: object = None

    def __call__(self, x):
        x = np.asarray(x, F)
        cin = x.shape[-1]
        kernel = self.param("kernel", None, (*self.kernel_size, cin, self.features))
        assert kernel.shape == (*self.kernel_size, cin, self.features), (self.name, kernel.shape)
        strides = tuple(self.strides) if self.strides is not None else (1, 1)
        if isinstance(self.padding, str):
            assert self.padding == "SAME"
            pads = [_same_pad(x.shape[-3], self.kernel_size[0], strides[0]), _same_pad(x.shape[-2], self.kernel_size[1], strides[1])]
        else:
            pads = [tuple(p) for p in self.padding]
        t = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        t = Fnn.pad(t, (pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
        w = torch.from_numpy(np.ascontiguousarray(kernel)).permute(3, 2, 0, 1)
        y = Fnn.conv2d(t, w, stride=strides).permute(0, 2, 3, 1).numpy()
        if self.use_bias:
            y = y + self.param("bias", None, (self.features,))
        return y.astype(F)


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None
    param_dtype: object = None
    dtype: object = None

    def __call__(self, x):
        y = np.asarray(x, F) @ self.param("kernel")
        if self.use_bias:
            y = y + self.param("bias")
        return y.astype(F)


def max_pool(x, window_shape, strides=None, padding="VALID"):
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, F))).permute(0, 3, 1, 2)
    pads = [tuple(p) for p in padding] if not isinstance(padding, str) else [(0, 0), (0, 0)]
    t = Fnn.pad(t, (pads[1][0], pads[1][1], pads[0][0], pads[0][1]), value=float("-inf"))
    return Fnn.max_pool2d(t, tuple(window_shape), stride=tuple(strides or window_shape)).permute(0, 2, 3, 1).numpy()


class Sequential(Module):
    layers: list

    def __call__(self, x):
        for i, layer in enumerate(self.layers):      # flax names the layers of a Sequential layers_<i> under it
            if isinstance(layer, Module) and layer._params is None:
                layer._parent = self
                object.__setattr__(layer, "name", f"layers_{i}")
        for layer in self.layers:
            x = layer(x) if not isinstance(x, tuple) else layer(*x)
        return x


def image_resize(x, shape, method):
    """jax.image.resize(x, shape, 'bilinear') for exact 2x up-sampling: half-pixel centres, edge clamp."""
    assert method == "bilinear"
    x = np.asarray(x, F)
    t = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
    y = Fnn.interpolate(t, size=tuple(shape[-3:-1]), mode="bilinear", align_corners=False)
    return y.permute(0, 2, 3, 1).numpy()


class ConfigDict(dict):
    """ml_collections.ConfigDict subset: attribute access, nested dicts converted, lock() / placeholder."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def lock(self):
        return self

    def unlocked(self):
        import contextlib
        return contextlib.nullcontext(self)


def install(mods, jax):
    """Replace the permissive flax stubs of jaxshim.install by this stand-in."""
    mlc = types.ModuleType("ml_collections")
    cd = types.ModuleType("ml_collections.config_dict")
    cd.ConfigDict, cd.placeholder = ConfigDict, (lambda t: None)
    mlc.ConfigDict, mlc.config_dict = ConfigDict, cd
    mods["ml_collections"], mods["ml_collections.config_dict"] = mlc, cd
    nn = types.ModuleType("flax.linen")
    for k, v in dict(Module=Module, compact=compact, remat=remat, relu=relu, Conv=Conv, Dense=Dense, max_pool=max_pool,
                     Sequential=Sequential).items():
        setattr(nn, k, v)
    nn.initializers = types.SimpleNamespace(ones=None, zeros=None, lecun_normal=lambda *a, **k: None, constant=lambda v: None)
    nn.log_sigmoid = lambda x: (-(np.maximum(-np.asarray(x), 0) + np.log1p(np.exp(-np.abs(np.asarray(x)))))).astype(F)
    nn.vmap = lambda cls, **kw: cls
    flax = mods["flax"]
    flax.linen = nn
    mods["flax.linen"] = nn
    jax.image.resize = image_resize
    jax.nn.initializers.lecun_normal = lambda *a, **k: None
    jax.nn.initializers.glorot_uniform = lambda *a, **k: None
    jax.nn.initializers.zeros = None
    return nn
