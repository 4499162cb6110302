"""Minimal torch-CPU stand-in for the `tensorflow` symbols keras_nerf's hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow cannot be installed offline, so
`tests/golden/make_golden.py` puts this directory first on sys.path and imports the reference's
own, unmodified Python (`/root/reference/keras_nerf/...`).  Each symbol restates the documented TF
semantics [TF-sem]; device-dependent behaviour is explicit:

* `tf.gather` with an out-of-range index: `config.gather_oob = "zero"` (TF-GPU kernel, default) or
  "raise" (TF-CPU kernel).
* `tf.random.uniform`: draws are popped from `random.queue` (a list of arrays pushed by the caller)
  so the same numbers can be fed to the oracle and the CUDA path; falls back to a seeded generator.
* scans (`cumsum`, `math.cumprod`) are sequential fp32 like TF's CPU kernels.

It is NOT a general TF emulation: only what rays.py, data/utils.py, model/nerf/{utils,mlp,nerf}.py use.
"""
from __future__ import annotations

import math as _math
from types import SimpleNamespace as _NS

import numpy as _np
import torch as _torch

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
newaxis = None

config = _NS(gather_oob="zero")


class Tensor(_torch.Tensor):
    """torch tensor with TF's immutable-value semantics for augmented assignment."""

    def __iadd__(self, other):  # `weights += 1e-5` must not mutate the caller's tensor
        return _torch.add(self, _c(other))

    def __isub__(self, other):
        return _torch.sub(self, _c(other))

    def __imul__(self, other):
        return _torch.mul(self, _c(other))

    def __itruediv__(self, other):
        return _torch.div(self, _c(other))

    def numpy(self):
        return self.detach().as_subclass(_torch.Tensor).numpy()

    def __eq__(self, other):  # used as `shape == (..)` only on .shape; keep tensor eq elementwise
        return _torch.Tensor.__eq__(self, other)

    __hash__ = _torch.Tensor.__hash__


class Variable:
    def __init__(self, initial_value, trainable=True, dtype=None, name=None):
        v = _c(initial_value, dtype).detach().clone().as_subclass(_torch.Tensor)
        self.trainable = trainable
        self.value = v.requires_grad_(bool(trainable) and v.is_floating_point())
        self.name = name

    @property
    def shape(self):
        return self.value.shape

    @property
    def dtype(self):
        return self.value.dtype

    def _set(self, new):
        new = _c(new).detach().clone().as_subclass(_torch.Tensor).to(self.value.dtype)
        self.value = new.requires_grad_(bool(self.trainable) and new.is_floating_point())
        return self

    def assign(self, v):
        return self._set(v)

    def assign_add(self, v):
        return self._set(self.value.detach() + _c(v).detach())

    def assign_sub(self, v):
        return self._set(self.value.detach() - _c(v).detach())

    def numpy(self):
        return self.value.detach().numpy()

    def __eq__(self, other):
        return bool((self.value == _c(other)).all())

    __hash__ = object.__hash__


def _c(x, dtype=None):
    """convert_to_tensor: python scalars/lists -> float32 (ints stay int32), Variables -> value."""
    if isinstance(x, Variable):
        t = x.value
    elif _torch.is_tensor(x):
        t = x
    elif isinstance(x, _np.ndarray):
        t = _torch.from_numpy(_np.ascontiguousarray(x))
        if t.dtype == _torch.float64 and dtype is None:
            t = t.to(float32)
    elif isinstance(x, (list, tuple)):
        if any(_torch.is_tensor(e) or isinstance(e, (list, tuple, Variable)) for e in x):
            t = _torch.stack([_c(e, dtype) for e in x])
        else:
            isint = all(isinstance(e, (int, _np.integer)) and not isinstance(e, bool) for e in x)
            t = _torch.tensor(x, dtype=int32 if isint else float32)
    elif isinstance(x, bool):
        t = _torch.tensor(x)
    elif isinstance(x, (int, _np.integer)):
        t = _torch.tensor(int(x), dtype=int32)
    else:
        t = _torch.tensor(float(x), dtype=float32)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.as_subclass(Tensor)


def _like(x, ref):
    """python scalars adopt the dtype of the tensor operand (TF binary-op conversion)."""
    if _torch.is_tensor(x) or isinstance(x, Variable):
        return _c(x)
    return _c(x, ref.dtype)


def _ints(shape):
    if _torch.is_tensor(shape):
        return [int(s) for s in shape.reshape(-1)]
    return [int(s) for s in shape]


def function(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def constant(value, dtype=None, shape=None):
    return _c(value, dtype)


convert_to_tensor = constant


def cast(x, dtype):
    return _c(x).to(dtype).as_subclass(Tensor)


def shape(x):
    return tuple(_c(x).shape)


def reshape(x, shp):
    return _c(x).reshape(_ints(shp))


def range(*args, dtype=None):  # noqa: A001
    a = [int(v) if not isinstance(v, float) else v for v in args]
    return _torch.arange(*a, dtype=dtype or int32).as_subclass(Tensor)


def meshgrid(a, b, indexing="xy"):
    x, y = _torch.meshgrid(_c(a), _c(b), indexing=indexing)
    return x.as_subclass(Tensor), y.as_subclass(Tensor)


def stack(values, axis=0):
    return _torch.stack([_c(v) for v in values], dim=axis).as_subclass(Tensor)


def concat(values, axis):
    return _torch.cat([_c(v) for v in values], dim=axis).as_subclass(Tensor)


def ones_like(x):
    return _torch.ones_like(_c(x))


def zeros_like(x):
    return _torch.zeros_like(_c(x).detach())


def broadcast_to(x, shape):  # noqa: A002
    return _c(x).expand(_ints(shape))


def reduce_sum(x, axis=None, keepdims=False):
    """[TF-sem] TF does not define a summation order (Eigen tree on CPU, cub on GPU); the stand-in sums
    strictly left to right in fp32 along `axis` so that the order is a documented, reproducible choice."""
    x = _c(x)
    if axis is None:
        return x.sum()
    xm = x.movedim(axis, -1)
    acc = xm[..., 0]
    for i in _builtin_range(1, xm.shape[-1]):
        acc = acc + xm[..., i]
    if keepdims:
        acc = acc.unsqueeze(axis)
    return acc.as_subclass(Tensor)


def reduce_mean(x, axis=None, keepdims=False):
    x = _c(x)
    if axis is None:
        return x.mean()
    return x.mean(dim=axis, keepdim=keepdims)


def norm(x, axis=None, keepdims=False):
    x = _c(x)
    return _torch.sqrt((x * x).sum(dim=axis, keepdim=keepdims))


def linspace(start, stop, num):
    """[TF-sem] start + ((stop-start)/(num-1))*i with exact end points."""
    s, e, n = _c(start, float32), _c(stop, float32), int(num)
    if n == 1:
        return s.reshape(1)
    delta = (e - s) / _torch.tensor(float(n - 1), dtype=float32)
    out = s + delta * _torch.arange(n, dtype=float32)
    out[0] = s
    out[-1] = e
    return out.as_subclass(Tensor)


def clip_by_value(x, lo, hi):
    x = _c(x)
    return _torch.clamp(x, min=float(_c(lo)), max=float(_c(hi)))


def exp(x):
    return _torch.exp(_c(x))


def sin(x):
    return _torch.sin(_c(x))


def cos(x):
    return _torch.cos(_c(x))


def tan(x):
    return _torch.tan(_c(x))


def maximum(a, b):
    ref = a if _torch.is_tensor(a) else _c(b)
    return _torch.maximum(_like(a, ref), _like(b, ref))


def minimum(a, b):
    ref = a if _torch.is_tensor(a) else _c(b)
    return _torch.minimum(_like(a, ref), _like(b, ref))


def where(cond, a, b):
    return _torch.where(cond, _c(a), _c(b))


def sort(x, axis=-1, direction="ASCENDING"):
    v, _ = _torch.sort(_c(x), dim=axis, descending=(direction != "ASCENDING"))
    return v


def searchsorted(sorted_sequence, values, side="left", out_type=int32):
    r = _torch.searchsorted(_c(sorted_sequence).contiguous(), _c(values).contiguous(),
                            right=(side == "right"))
    return r.to(out_type).as_subclass(Tensor)


def cumsum(x, axis=-1, exclusive=False, reverse=False):
    assert not exclusive and not reverse
    x = _c(x).movedim(axis, -1)
    outs, acc = [], None
    for i in _builtin_range(x.shape[-1]):
        acc = x[..., i] if acc is None else acc + x[..., i]
        outs.append(acc)
    return _torch.stack(outs, dim=-1).movedim(-1, axis).as_subclass(Tensor)


def _cumprod(x, axis=-1, exclusive=False, reverse=False):
    assert not reverse
    x = _c(x).movedim(axis, -1)
    outs = []
    acc = _torch.ones_like(x[..., 0])
    for i in _builtin_range(x.shape[-1]):
        if exclusive:
            outs.append(acc)
            acc = acc * x[..., i]
        else:
            acc = acc * x[..., i]
            outs.append(acc)
    return _torch.stack(outs, dim=-1).movedim(-1, axis).as_subclass(Tensor)


import builtins as _builtins  # noqa: E402

_builtin_range = _builtins.range


def gather(params, indices, axis=-1, batch_dims=0):
    """tf.gather(params[..., N], indices[..., K, 2], axis=-1, batch_dims=rank-2) as used at
    model/nerf/utils.py:83-88.  Out-of-range indices follow config.gather_oob [TF-sem]."""
    params, indices = _c(params), _c(indices).to(int64)
    assert axis in (-1, params.dim() - 1)
    assert batch_dims == params.dim() - 1, "only the batched last-axis gather of the hot path"
    n = params.shape[-1]
    flat = indices.reshape(indices.shape[:batch_dims] + (-1,))
    bad = (flat < 0) | (flat >= n)
    if bool(bad.any()):
        if config.gather_oob == "raise":
            raise IndexError(f"InvalidArgumentError: indices out of range [0, {n}) (TF-CPU gather)")
        assert config.gather_oob == "zero"
    safe = flat.clamp(0, n - 1)
    out = _torch.gather(params, -1, safe)
    out = _torch.where(bad, _torch.zeros_like(out), out)
    return out.reshape(indices.shape).as_subclass(Tensor)


def print(*a, **k):  # noqa: A001  (tf.print)
    pass


# ---- tf.random -------------------------------------------------------------------------------
def _uniform(shape, minval=0, maxval=None, dtype=float32, seed=None):
    shp = _ints(shape)
    if random.queue:
        arr = _np.asarray(random.queue.pop(0), dtype=_np.float32)
        assert int(arr.size) == int(_np.prod(shp)), (arr.shape, shp)
        return _torch.from_numpy(arr.reshape(shp).copy()).as_subclass(Tensor)
    return _torch.rand(shp, generator=random.generator, dtype=float32).as_subclass(Tensor)


def _set_seed(seed):
    random.generator = _torch.Generator().manual_seed(int(seed))


random = _NS(uniform=_uniform, set_seed=_set_seed, queue=[], generator=_torch.Generator().manual_seed(0))

math = _NS(cumprod=_cumprod, cumsum=cumsum,
           count_nonzero=lambda x: (_c(x) != 0).sum().to(int64))


# ---- tf.debugging ----------------------------------------------------------------------------
def _assert_all_finite(x, message):
    if not bool(_torch.isfinite(_c(x)).all()):
        raise FloatingPointError(message)
    return x


debugging = _NS(assert_all_finite=_assert_all_finite)


# ---- tf.image --------------------------------------------------------------------------------
def _psnr(a, b, max_val):
    a, b = _c(a), _c(b)
    m = ((a - b) ** 2).reshape(a.shape[0], -1).mean(dim=1)
    return (20.0 * _math.log10(max_val) - 10.0 * _torch.log10(m)).as_subclass(Tensor)


def _ssim(a, b, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03):
    """tf.image.ssim [TF-sem]: 11x11 gaussian window (sigma 1.5), VALID, per channel, mean -> [B]."""
    a, b = _c(a).permute(0, 3, 1, 2), _c(b).permute(0, 3, 1, 2)
    C = a.shape[1]
    g = _torch.arange(filter_size, dtype=float32) - (filter_size - 1) / 2.0
    g = _torch.exp(-(g * g) / (2.0 * filter_sigma * filter_sigma))
    g = g / g.sum()
    k = (g[:, None] * g[None, :]).expand(C, 1, filter_size, filter_size).contiguous()
    conv = lambda x: _torch.nn.functional.conv2d(x, k, groups=C)  # noqa: E731
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    mu_a, mu_b = conv(a), conv(b)
    s_aa, s_bb, s_ab = conv(a * a) - mu_a * mu_a, conv(b * b) - mu_b * mu_b, conv(a * b) - mu_a * mu_b
    lum = (2 * mu_a * mu_b + c1) / (mu_a * mu_a + mu_b * mu_b + c1)
    cs = (2 * s_ab + c2) / (s_aa + s_bb + c2)
    return (lum * cs).mean(dim=(1, 2, 3)).as_subclass(Tensor)


image = _NS(psnr=_psnr, ssim=_ssim)


# ---- GradientTape / TensorArray --------------------------------------------------------------
class GradientTape:
    def __init__(self, watch_accessed_variables=True, persistent=False):
        self._watched = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, variables):
        self._watched.extend(variables if isinstance(variables, (list, tuple)) else [variables])

    def gradient(self, target, sources):
        leaves = [v.value for v in sources]
        grads = _torch.autograd.grad(_c(target), leaves, allow_unused=True, retain_graph=False)
        return [(g if g is not None else _torch.zeros_like(l)).as_subclass(Tensor)
                for g, l in zip(grads, leaves)]


class TensorArray:
    def __init__(self, dtype, size=0, **kwargs):
        self._items = [None] * int(size)

    def write(self, i, value):
        self._items[int(i)] = _c(value)
        return self

    def stack(self):
        return _torch.stack(self._items, dim=0).as_subclass(Tensor)


# ---- tf.keras --------------------------------------------------------------------------------
class _Dense:
    def __init__(self, units, activation=None, name=None, kernel_initializer="glorot_uniform", **kw):
        assert kernel_initializer == "glorot_uniform"
        self.units, self.activation, self.name = int(units), activation, name
        self.kernel = self.bias = None

    def build(self, fan_in):
        lim = _math.sqrt(6.0 / (fan_in + self.units))               # glorot_uniform [TF-sem]
        w = keras.init_rng.uniform(-lim, lim, size=(fan_in, self.units)).astype(_np.float32)
        self.kernel = Variable(_torch.from_numpy(w), name=f"{self.name}/kernel:0")
        self.bias = Variable(_torch.zeros(self.units, dtype=float32), name=f"{self.name}/bias:0")

    def __call__(self, x):
        x = _c(x)
        if self.kernel is None:
            self.build(int(x.shape[-1]))
        y = _torch.matmul(x, self.kernel.value) + self.bias.value
        if self.activation == "relu":
            y = _torch.relu(y)
        elif self.activation == "sigmoid":
            y = _torch.sigmoid(y)
        else:
            assert self.activation is None, self.activation
        return y.as_subclass(Tensor)

    @property
    def trainable_variables(self):
        return [] if self.kernel is None else [self.kernel, self.bias]


class _Model:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self.run_eagerly = False

    def compile(self, run_eagerly=False, **kwargs):
        self.run_eagerly = bool(run_eagerly)

    def __call__(self, inputs, **kwargs):
        return self.call(inputs)

    def get_config(self):
        return {"name": self.name}

    @property
    def trainable_variables(self):
        out = []
        for v in vars(self).values():
            items = v if isinstance(v, (list, tuple)) else [v]
            for it in items:
                if isinstance(it, (_Dense, _Model)):
                    out.extend(it.trainable_variables)
        return out

    def summary(self):
        pass

    def save_weights(self, path):
        _np.savez(path, *[v.numpy() for v in self.trainable_variables])

    def load_weights(self, path):
        import os
        p = path if os.path.exists(path) else path + ".npz"
        z = _np.load(p)
        for v, k in zip(self.trainable_variables, z.files):
            v.assign(z[k])


class _Adam:
    """Keras Adam defaults [TF-sem]: lr 1e-3, beta_1 .9, beta_2 .999, epsilon 1e-7."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._m, self._v = {}, {}

    def apply_gradients(self, grads_and_vars):
        self.iterations += 1
        t = self.iterations
        lr_t = _np.float32(self.lr * _math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t))
        for g, var in grads_and_vars:
            g = _c(g).detach().as_subclass(_torch.Tensor)
            k = id(var)
            if k not in self._m:
                self._m[k] = _torch.zeros_like(g)
                self._v[k] = _torch.zeros_like(g)
            self._m[k] = self._m[k] + (g - self._m[k]) * _np.float32(1.0 - self.b1)
            self._v[k] = self._v[k] + (g * g - self._v[k]) * _np.float32(1.0 - self.b2)
            var.assign_sub(lr_t * self._m[k] / (_torch.sqrt(self._v[k]) + _np.float32(self.eps)))


def _get_optimizer(identifier):
    if isinstance(identifier, str):
        assert identifier.lower() == "adam", identifier
        return _Adam()
    return identifier


class _Mean:
    def __init__(self, name=None):
        self.name, self.total, self.count = name, 0.0, 0

    def update_state(self, values):
        v = _c(values).detach().reshape(-1).to(float32)
        self.total += float(v.sum())
        self.count += int(v.numel())

    def result(self):
        return _torch.tensor(self.total / max(self.count, 1), dtype=float32).as_subclass(Tensor)

    def reset_state(self):
        self.total, self.count = 0.0, 0


class _MSE:
    def __init__(self, reduction="auto"):
        self.reduction = reduction

    def __call__(self, y_true, y_pred):
        d = _c(y_pred) - _c(y_true)
        per = (d * d).mean(dim=-1)
        if self.reduction in ("none", None):
            return per
        return per.mean()


keras = _NS(
    Model=_Model,
    layers=_NS(Dense=_Dense,
               concatenate=lambda xs, axis=-1: _torch.cat([_c(x) for x in xs], dim=axis).as_subclass(Tensor)),
    optimizers=_NS(get=_get_optimizer, Adam=_Adam),
    metrics=_NS(Mean=_Mean),
    losses=_NS(MeanSquaredError=_MSE, Reduction=_NS(NONE="none")),
    init_rng=_np.random.default_rng(42),
)
