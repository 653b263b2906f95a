"""TEST INFRASTRUCTURE -- a torch-backed stand-in for the slice of the TensorFlow 2.8 / Keras API that the
reference's MFP hot path touches, so that the reference's OWN Python (``/root/reference/src/mfp/mfp``) can be
imported and executed in this container (TensorFlow itself is not installable here: no wheel, Python 3.12).

Used by ``tests/golden/make_golden.py`` ONLY (it puts this directory on ``sys.path`` ahead of everything else and
then imports ``mfp.models.mfp`` from the read-only reference tree).  Nothing in the product imports it and it does
not travel into any measured path.

What this does and does not pin: every line of the reference's Python on the path (masking variants and their
per-document selection, encoder fusion, block wiring, heads, LossLayer weighting / sorting, merge) runs unmodified.
The semantics of the TF/Keras primitives themselves (LayerNormalization epsilon 1e-3, Keras
CE clip 1e-7, L2 = l2*sum(w^2), inverted dropout, per-variable clipnorm, TF Adam epsilon placement ...) are restated
here from recall, exactly as in ``oracle/mfp_oracle.py`` -- they stay unverifiable offline.

Tensors are plain ``torch.Tensor``.  ``tf.float32`` maps to ``config.FLOAT`` (float32 for bit-faithful masking runs,
float64 for high-precision network/loss goldens).  Random draws come from ``config.rng`` (scripted by the caller).
"""
import builtins as _b
import sys
import types

import numpy as np
import torch

from . import config

class TensorShape(tuple):
    pass


class Tensor(torch.Tensor):
    """TF tensors are immutable: ``x += y`` rebinds the name.  The reference relies on that (mask.py:29 ``length += 1``,
    transformer.py:217 ``x += y`` on the block input), so augmented assignment is out-of-place here as well."""

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def __itruediv__(self, o):
        return self / o

    def __iand__(self, o):
        return self & o

    def __ior__(self, o):
        return self | o

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()


def _wrap(x):
    if isinstance(x, torch.Tensor) and not isinstance(x, Tensor):
        return x.as_subclass(Tensor)
    if isinstance(x, (list, tuple)) and not isinstance(x, TensorShape) and any(isinstance(v, torch.Tensor) for v in x):
        return type(x)(_wrap(v) for v in x)
    return x


def _wrapping(fn):
    def inner(*a, **k):
        return _wrap(fn(*a, **k))

    inner.__name__ = getattr(fn, "__name__", "fn")
    return inner


newaxis = None
bool = torch.bool  # noqa: A001
int32 = torch.int32
int64 = torch.int64


class _FloatAlias:
    """tf.float32 -> config.FLOAT, resolved at use time."""

    def __repr__(self):
        return "tf.float32(stand-in:%s)" % config.FLOAT


float32 = _FloatAlias()


def _dt(dtype):
    return config.FLOAT if isinstance(dtype, _FloatAlias) else dtype


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(_dt(dtype))
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(x)
    else:
        t = torch.as_tensor(x)
    if t.dtype in (torch.float32, torch.float64) and dtype is None:
        t = t.to(config.FLOAT)
    return t if dtype is None else t.to(_dt(dtype))


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def executing_eagerly():
    return True


def function(*a, **k):
    if a and callable(a[0]):
        return a[0]
    return lambda f: f


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def identity(x):
    return x.clone() if isinstance(x, torch.Tensor) else x


def stop_gradient(x):
    return x.detach()


def shape(x):
    return tuple(int(s) for s in _t(x).shape)


def rank(x):
    return _Rank(_t(x).dim())


class _Rank(int):
    def numpy(self):
        return int(self)


def size(x):
    return _t(x).numel()


def cast(x, dtype):
    d = _dt(dtype)
    x = _t(x)
    if d in (torch.int32, torch.int64) and x.is_floating_point():
        return torch.trunc(x).to(d)
    return x.to(d)


def reshape(x, shp):
    return _t(x).reshape(tuple(int(s) for s in shp))


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0):
    return torch.stack(list(values), dim=axis)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def repeat(x, repeats, axis=None):
    return torch.repeat_interleave(x, int(repeats), dim=axis)


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def squeeze(x, axis=None):
    return x.squeeze() if axis is None else x.squeeze(axis)


def transpose(x, perm=None):
    return x.permute(*perm) if perm is not None else x.T


def split(x, num, axis=0):
    return list(torch.chunk(x, num, dim=axis))


def range(*args):  # noqa: A001
    return torch.arange(*[int(a) for a in args], dtype=torch.int32)


def fill(dims, value):
    dims = tuple(int(d) for d in dims)
    if isinstance(value, _b.bool):
        return torch.full(dims, value, dtype=torch.bool)
    return torch.full(dims, value, dtype=config.FLOAT if isinstance(value, float) else torch.int32)


def zeros(shp, dtype=float32):
    shp = (int(shp),) if isinstance(shp, int) else tuple(int(s) for s in shp)
    return torch.zeros(shp, dtype=_dt(dtype))


def ones(shp, dtype=float32):
    shp = (int(shp),) if isinstance(shp, int) else tuple(int(s) for s in shp)
    return torch.ones(shp, dtype=_dt(dtype))


def zeros_like(x):
    return torch.zeros_like(x)


def ones_like(x):
    return torch.ones_like(x)


def eye(n):
    return torch.eye(int(n), dtype=config.FLOAT)


def one_hot(indices, depth):
    return torch.nn.functional.one_hot(_t(indices).to(torch.int64), int(depth)).to(config.FLOAT)


def where(cond, x=None, y=None):
    if not isinstance(x, torch.Tensor) and not isinstance(y, torch.Tensor):
        x = _t(x)
    if isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) and x.dtype != y.dtype:
        raise TypeError("tf.where: dtype mismatch %s vs %s" % (x.dtype, y.dtype))  # TF is strict about this
    return torch.where(cond, x, y)


def gather(params, indices, axis=None, batch_dims=0):
    params = _t(params)
    idx = _t(indices).to(torch.int64)
    if batch_dims == 0:
        assert axis in (None, 0)
        return params[idx]
    assert batch_dims == 1 and idx.dim() == 2
    full = idx.reshape(idx.shape + (1,) * (params.dim() - 2)).expand(-1, -1, *params.shape[2:])
    return torch.gather(params, 1, full)


def argmax(x, axis=None, output_type=int64):
    return torch.argmax(x, dim=axis).to(output_type)


def argsort(x, axis=-1, direction="ASCENDING", stable=False):
    # TF's argsort(stable=False) gives no order guarantee among equal keys; a stable sort is one valid outcome.
    return torch.argsort(x, dim=axis, stable=True, descending=(direction != "ASCENDING")).to(torch.int32)


def sort(x, axis=-1, direction="ASCENDING"):
    return torch.sort(x, dim=axis, descending=(direction != "ASCENDING")).values


def sequence_mask(lengths, maxlen=None):
    lengths = _t(lengths).to(torch.int64)
    n = int(lengths.max()) if maxlen is None else int(maxlen)
    return torch.arange(n)[None, :] < lengths[..., None]


def reduce_sum(x, axis=None, keepdims=False):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None):
    return x.max() if axis is None else x.max(dim=axis).values


def reduce_min(x, axis=None):
    return x.min() if axis is None else x.min(dim=axis).values


def matmul(a, b, transpose_a=False, transpose_b=False):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def logical_not(x):
    return ~x


def logical_or(a, b):
    return a | b


def logical_and(a, b):
    return a & b


def minimum(a, b):
    return torch.minimum(_t(a), _t(b))


def meshgrid(*a, **k):
    return torch.meshgrid(*a, indexing=k.get("indexing", "xy"))


def assert_rank(x, r, *a, **k):
    assert _t(x).dim() == r, ("assert_rank", tuple(_t(x).shape), r)


# ------------------------------------------------------------------------------------------------ submodules
def _module(name, **members):
    m = types.ModuleType(name)
    m.__dict__.update(members)
    sys.modules[name] = m
    return m


def _assert_equal(a, b, *args, **k):
    assert torch.equal(_t(a), _t(b)) if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor) else a == b, ("assert_equal", a, b)


def _assert_rank_at_least(x, r, *a, **k):
    assert _t(x).dim() >= r


def _assert_less_equal(a, b, *args, **k):
    assert _b.bool(torch.all(_t(a) <= _t(b))), ("assert_less_equal", a, b)


def _assert_greater_equal(a, b, *args, **k):
    assert _b.bool(torch.all(_t(a) >= _t(b))), ("assert_greater_equal", a, b)


debugging = _module(__name__ + ".debugging", assert_rank=assert_rank, assert_equal=_assert_equal,
                    assert_rank_at_least=_assert_rank_at_least, assert_less_equal=_assert_less_equal,
                    assert_greater_equal=_assert_greater_equal)


def _log(x):
    return torch.log(_t(x))


math = _module(__name__ + ".math", log=_log, sqrt=lambda x: torch.sqrt(_t(x)), abs=lambda x: torch.abs(x),
               reduce_all=lambda x, axis=None: x.all() if axis is None else x.all(dim=axis),
               logical_or=logical_or, logical_and=logical_and, logical_not=logical_not,
               minimum=minimum, maximum=lambda a, b: torch.maximum(_t(a), _t(b)),
               truediv=lambda a, b: _t(a) / _t(b), is_finite=lambda x: torch.isfinite(x),
               argmax=argmax)

nn = _module(__name__ + ".nn", softmax=lambda x, axis=-1: torch.softmax(x, dim=axis), relu=torch.relu)


def _band_part(x, lower, upper):
    n, m = x.shape[-2:]
    i = torch.arange(n)[:, None]
    j = torch.arange(m)[None, :]
    keep = torch.ones(n, m, dtype=torch.bool)
    if lower >= 0:
        keep &= (i - j) <= lower
    if upper >= 0:
        keep &= (j - i) <= upper
    return x * keep.to(x.dtype)


linalg = _module(__name__ + ".linalg", band_part=_band_part, diag_part=lambda x: torch.diagonal(x, dim1=-2, dim2=-1))


def _uniform(shp, minval=0, maxval=None, dtype=float32, seed=None):
    shp = tuple(int(s) for s in shp)
    d = _dt(dtype)
    if d in (torch.int32, torch.int64):
        return _t(config.rng.randint(shp, int(minval), int(maxval))).to(d)
    assert minval in (0, 0.0) and maxval in (None, 1, 1.0)
    return _t(config.rng.uniform(shp)).to(d)


def _normal(shp, mean=0.0, stddev=1.0, dtype=float32, seed=None):
    shp = tuple(int(s) for s in shp)
    return _t(config.rng.normal(shp, float(stddev))).to(_dt(dtype))


random = _module(__name__ + ".random", uniform=_uniform, normal=_normal)

autograph = _module(__name__ + ".autograph")
autograph.experimental = _module(__name__ + ".autograph.experimental", set_loop_options=lambda **k: None)


def random_normal_initializer(**k):
    raise NotImplementedError("not on the MFP hot path")


def zeros_initializer(**k):
    raise NotImplementedError("not on the MFP hot path")


def Variable(*a, **k):
    raise NotImplementedError("not on the MFP hot path")


def TensorArray(*a, **k):
    raise NotImplementedError("not on the MFP hot path")


def scatter_nd(*a, **k):
    raise NotImplementedError("not on the MFP hot path")


def tensor_scatter_nd_update(*a, **k):
    raise NotImplementedError("not on the MFP hot path")



def _wrap_module(mod):
    for _name, _val in list(vars(mod).items()):
        if _name.startswith("_") or not isinstance(_val, types.FunctionType) and not isinstance(_val, type(torch.relu)):
            continue
        if _name in ("shape", "rank", "size", "function", "is_tensor", "executing_eagerly"):
            continue
        setattr(mod, _name, _wrapping(_val))


for _m in (sys.modules[__name__], math, nn, linalg, random):
    _wrap_module(_m)

from . import keras  # noqa: E402,F401

# import-time references of code that is NOT on the hot path (data/spec.py:219 default argument, etc.)
data = _module(__name__ + ".data")
data.experimental = _module(__name__ + ".data.experimental", AUTOTUNE=-1)
io = _module(__name__ + ".io")
__version__ = "2.8.0"  # README.md:10 of the reference; selects the non-2.3/2.4 branches of data/spec.py and data/discretizer.py

# input side (data/spec.py, data/discretizer.py): see data_io.py
from . import data_io as _data_io  # noqa: E402

io.FixedLenFeature = _data_io.FixedLenFeature
io.FixedLenSequenceFeature = _data_io.FixedLenSequenceFeature
io.parse_sequence_example = _data_io.parse_sequence_example
io.gfile = _module(__name__ + ".io.gfile", GFile=_data_io._GFile)
_experimental = keras._mod(keras.__name__ + ".layers.experimental")
_experimental.preprocessing = keras._mod(keras.__name__ + ".layers.experimental.preprocessing", StringLookup=_data_io.StringLookup,
                                         IntegerLookup=_data_io.IntegerLookup, Discretization=_data_io.Discretization)
keras.layers.experimental = _experimental

# einops (encoder.py uses rearrange on the context token) picks its backend by the tensor's type and would try a TensorFlow backend
# because a module named "tensorflow" is loaded: register the torch backend first -- stand-in tensors are torch tensors.
try:
    from einops import _backends as _einops_backends

    _einops_backends._loaded_backends.setdefault("torch", _einops_backends.TorchBackend())
except Exception:  # einops absent: nothing on the path needs it then
    pass
