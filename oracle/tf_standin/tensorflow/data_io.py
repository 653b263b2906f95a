"""Input-side slice of the TensorFlow stand-in (TEST INFRASTRUCTURE; used by tests/golden/make_dataspec_golden.py only): what the
reference's ``data/spec.py`` and ``data/discretizer.py`` import -- ``tf.io.gfile.GFile``, ``tf.io.parse_sequence_example`` with
``FixedLenFeature`` / ``FixedLenSequenceFeature``, and the Keras preprocessing layers ``StringLookup`` / ``IntegerLookup`` /
``Discretization`` -- so that the reference's own ``DataSpec`` (schema files included) runs unmodified here.

The record decoding and the layer semantics come from ``oracle/dataspec_oracle.py`` (documented TensorFlow behaviour restated; see its
header).  What a golden made this way pins: the reference's YAML specs, ``_create_lookup`` (vocabulary filtering by ``min_freq``,
integer ranges, option handling), ``make_input_columns`` (``input_dim``, ``primary_label``, ``loss_condition`` masks), ``parse_fn``
(feature specs from the columns, preprocessing order, int64 -> int32 cast) and ``unbatch`` / ``logit_to_label``.
"""
import collections

import numpy as np
import torch

from oracle import dataspec_oracle as DO

FixedLenFeature = collections.namedtuple("FixedLenFeature", ["shape", "dtype", "default_value"], defaults=[None])
FixedLenSequenceFeature = collections.namedtuple("FixedLenSequenceFeature", ["shape", "dtype", "allow_missing", "default_value"], defaults=[False, None])


class StringTensor(np.ndarray):
    """A tf.string tensor: numpy object array of bytes with the two Tensor methods the reference calls on it."""

    def numpy(self):
        return np.asarray(self)


def _strings(values, shape):
    arr = np.empty(len(values), dtype=object)
    arr[:] = values
    return arr.reshape(shape).view(StringTensor)


def _kind(dtype):
    return {"int64": "int64", "float32": "float32", "string": "string"}[str(dtype)]


def parse_sequence_example(serialized, context_features=None, sequence_features=None, **_):
    """tf.io.parse_sequence_example: dense context features, dense sequence features padded to the batch maximum."""
    decoded = [DO.decode_sequence_example(s) for s in serialized]
    B = len(decoded)
    default = {"int64": 0, "float32": 0.0, "string": b""}
    context, sequence, lengths = {}, {}, {}

    def tensor(flat, shape, kind):
        if kind == "string":
            return _strings(flat, shape)
        return torch.tensor(np.asarray(flat, dtype=kind).reshape(shape))

    for name, spec in (context_features or {}).items():
        kind, width = _kind(spec.dtype), int(np.prod(spec.shape))
        flat = []
        for ctx, _lists in decoded:
            assert name in ctx, "Feature %s is required but could not be found" % name
            k, values = ctx[name]
            assert k == kind and len(values) == width, name
            flat += values
        context[name] = tensor(flat, (B,) + tuple(spec.shape), kind)
    for name, spec in (sequence_features or {}).items():
        kind, width = _kind(spec.dtype), int(np.prod(spec.shape))
        steps = []
        for _ctx, lists in decoded:
            assert name in lists, "Feature list %s is required but could not be found" % name
            for k, values in lists[name]:
                assert k == kind and len(values) == width, name
            steps.append([v for _, v in lists[name]])
        T = max((len(s) for s in steps), default=0)
        flat = []
        for s in steps:
            for t in range(T):
                flat += s[t] if t < len(s) else [default[kind]] * width
        sequence[name] = tensor(flat, (B, T) + tuple(spec.shape), kind)
        lengths[name] = torch.tensor([len(s) for s in steps], dtype=torch.int64)
    return context, sequence, lengths


class _GFile:
    def __init__(self, path, mode="r"):
        self._f = open(path, mode)

    def __enter__(self):
        return self._f

    def __exit__(self, *exc):
        self._f.close()


class _IndexLookup:
    """[recall] Keras IndexLookup in "int" mode: tokens = [mask_token] + [oov_token] * num_oov_indices + vocabulary."""

    oov_token = None
    is_string = False

    def __init__(self, vocabulary=None, num_oov_indices=1, mask_token=None, **kwargs):
        mask_token = kwargs.pop("mask_value", mask_token)  # TF <= 2.5 spelling used by the reference's spec files
        assert not kwargs, kwargs
        column = {"dtype": "string" if self.is_string else "int64", "lookup": {"num_oov_indices": num_oov_indices, "mask_token": mask_token}}
        self._impl = DO.Lookup(column, "x", {"x": list(vocabulary)})

    def get_vocabulary(self):
        return list(self._impl.tokens)

    def vocabulary_size(self):
        return len(self._impl.tokens)

    def __call__(self, inputs):
        if isinstance(inputs, (str, bytes, int)):
            return torch.tensor(self._impl(inputs), dtype=torch.int64)
        arr = inputs.numpy() if hasattr(inputs, "numpy") else np.asarray(inputs)
        flat = [self._impl(v) for v in arr.reshape(-1).tolist()]
        return torch.tensor(flat, dtype=torch.int64).reshape(arr.shape)


class StringLookup(_IndexLookup):
    is_string = True


class IntegerLookup(_IndexLookup):
    pass


class Discretization:
    """[recall] Keras Discretization(bin_boundaries): Bucketize on float32 = number of boundaries <= x, int64 result."""

    def __init__(self, bin_boundaries=None, **kwargs):
        assert not kwargs, kwargs
        self.bin_boundaries = list(bin_boundaries)
        self._f32 = [np.float32(b) for b in self.bin_boundaries]

    def __call__(self, inputs):
        x = inputs.detach().cpu().numpy().astype(np.float32)
        out = np.zeros(x.shape, dtype=np.int64)
        for b in self._f32:
            out += (b <= x)
        return torch.tensor(out)
