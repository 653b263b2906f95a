"""``DataSpec``: the TFRecord side of the MFP step (reference: ``src/mfp/mfp/data/spec.py:25-362``, ``data/discretizer.py``).

Same constructor, properties and methods as the reference class -- ``DataSpec(name, path, batch_size)``, ``columns``, ``preprocessor``,
``size``, ``steps_per_epoch``, ``make_input_columns``, ``make_dataset``, ``parse_fn``, ``logit_to_label``, ``unbatch`` -- so ``train.py:38-52``
and ``eval.py:139-153`` read the same.  What TensorFlow does for the reference (``TFRecordDataset``, ``parse_sequence_example``,
``StringLookup`` / ``IntegerLookup`` / ``Discretization``) is done by ``libflexdm_io.so`` (``csrc/io/``; C ABI ``include/flexdm_io.h``):
records are mmapped, parsed on host threads and written, already looked-up / discretised / cast to int32, straight into pinned
batch buffers that ``DevicePrefetcher`` copies to the GPU.  There is no Python parsing path.

Directory layout read (spec.py:29-36): ``root/count.json``, ``root/vocabulary.json``, ``root/<split>-*.tfrecord``.
The column schemas of the two datasets (``data/crello-spec.yml``, ``data/rico-spec.yml``) are restated in ``BUILTIN_SPECS``; a path to
a YAML spec file of the same grammar is accepted as ``name`` too.
"""
import ctypes
import glob
import json
import os
import queue
import threading
from collections import OrderedDict
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import io_lib
from .spec import ATTRIBUTE_GROUPS  # noqa: F401  (the notebooks import it from mfp.data.spec, spec.py:364-377)


# ----------------------------------------------------------------------------------------------------------------------------------
# Column schemas (data/crello-spec.yml, data/rico-spec.yml), in file order: the order fixes the fusion-sum and head order.
def _seq(dtype, **kw):
    c = OrderedDict(is_sequence=True, dtype=dtype)
    c.update(kw)
    return c


def _unit_bins(bins):
    return {"min": 0.0, "max": 1.0, "bins": bins}


def _when_type(*values):
    return {"key": "type", "values": list(values)}


_LENGTH = OrderedDict(dtype="int64", lookup={"vocabulary": {"min": 1, "max": 50}, "num_oov_indices": 0, "mask_value": None})
_MASKED = {"mask_token": "", "num_oov_indices": 0}
_OOV = {"num_oov_indices": 1, "mask_token": None}


def _crello_columns():
    c = OrderedDict()
    c["id"] = OrderedDict(dtype="string", demo_only=True)
    c["length"] = _LENGTH
    c["group"] = OrderedDict(dtype="string", lookup=dict(_MASKED))
    c["format"] = OrderedDict(dtype="string", lookup=dict(_MASKED))
    c["canvas_width"] = OrderedDict(dtype="int64", lookup={"num_oov_indices": 0})
    c["canvas_height"] = OrderedDict(dtype="int64", lookup={"num_oov_indices": 0})
    c["category"] = OrderedDict(dtype="string", lookup=dict(_MASKED))
    c["type"] = _seq("string", lookup=dict(_MASKED), primary_label={"default": ""})
    for key in ("left", "top", "width", "height"):
        c[key] = _seq("float32", discretize=_unit_bins(64))
    c["opacity"] = _seq("float32", discretize=_unit_bins(8))
    c["color"] = _seq("int64", shape=[3], discretize={"min": 0, "max": 255, "bins": 16}, loss_condition=_when_type("textElement", "coloredBackground"))
    c["image_embedding"] = _seq("float32", shape=[512], loss_condition=_when_type("svgElement", "imageElement", "maskElement"))
    c["text_embedding"] = _seq("float32", shape=[512], loss_condition=_when_type("textElement"))
    c["font_family"] = _seq("string", min_freq=500, lookup=dict(_OOV), loss_condition=_when_type("textElement"))
    c["uuid"] = _seq("string", demo_only=True)
    return c


def _rico_columns():
    c = OrderedDict()
    c["length"] = _LENGTH
    for key in ("left", "top", "width", "height"):
        c[key] = _seq("float32", discretize=_unit_bins(64))
    c["clickable"] = _seq("int64", max=1)
    c["type"] = _seq("string", lookup=dict(_OOV), primary_label={"default": ""})
    c["icon"] = _seq("string", min_freq=500, lookup=dict(_OOV))
    c["text_button"] = _seq("string", min_freq=500, lookup=dict(_OOV))
    return c


BUILTIN_SPECS = {
    "crello": {"name": "crello", "columns": _crello_columns()},
    "rico": {"name": "rico", "columns": _rico_columns()},
}

_DTYPES = {"int": io_lib.INT64, "int32": io_lib.INT64, "int64": io_lib.INT64, "float": io_lib.FLOAT32, "float32": io_lib.FLOAT32,
           "float64": io_lib.FLOAT32, "string": io_lib.STRING}


# ----------------------------------------------------------------------------------------------------------------------------------
# Preprocessor objects: what DataSpec.preprocessor holds in the reference (Keras layers).  They carry the vocabulary / boundaries that
# the native parser applies; calling one directly (primary_label, spec.py:188-191, and un-preprocessing in unbatch) is host-side
# bookkeeping on a handful of values.
class _Lookup:
    """Keras ``StringLookup`` / ``IntegerLookup`` in ``output_mode="int"``: ``[mask_token] + [OOV] * num_oov_indices + vocabulary``.

    Version note.  The target is the API the reference states it was verified on, TensorFlow 2.8 (README.md:8-10), where both layers default
    to NO mask slot (``mask_token=None``).  The reference's spec files still carry the pre-2.6 keyword ``mask_value`` for ``length``
    (data/crello-spec.yml:6-13); it is honoured when given.  On TensorFlow <= 2.5 an ``IntegerLookup`` WITHOUT an explicit
    ``mask_value`` reserved index 0 (default ``mask_value=0``): for crello's ``canvas_width`` / ``canvas_height``
    (``lookup: {num_oov_indices: 0}``) such a checkpoint has one more row in those tables (``--context canvas`` / ``canvas_add`` only) and
    every index shifted by one, and does not load here (``load_weights`` reports the shape mismatch); pass ``mask_value: 0`` in the spec
    to reproduce that layout.  Not checkable offline: TensorFlow cannot be installed in this environment."""

    oov_token = None

    def __init__(self, vocabulary: Sequence, num_oov_indices: int = 1, mask_token=None, mask_value="__unset__", oov_token=None, **unknown):
        if unknown:
            raise TypeError("Unknown lookup options: %s" % sorted(unknown))
        if mask_value != "__unset__":  # TF <= 2.5 spelling, used by the reference's spec files for `length`
            mask_token = mask_value
        if num_oov_indices not in (0, 1):
            raise NotImplementedError("num_oov_indices > 1 (hashed OOV buckets) is not supported")
        self.vocabulary = [self._norm(v) for v in vocabulary]
        if len(set(self.vocabulary)) != len(self.vocabulary):
            raise ValueError("The passed vocabulary has repeated terms")
        self.num_oov_indices = int(num_oov_indices)
        self.mask_token = None if mask_token is None else self._norm(mask_token)
        if oov_token is not None:
            self.oov_token = self._norm(oov_token)
        if self.mask_token is not None and self.mask_token in self.vocabulary:
            raise ValueError("Reserved mask token %r found in the vocabulary" % (self.mask_token,))
        self._offset = (0 if self.mask_token is None else 1) + self.num_oov_indices
        self._index = {v: i + self._offset for i, v in enumerate(self.vocabulary)}

    def get_vocabulary(self) -> List:
        head = ([] if self.mask_token is None else [self.mask_token]) + [self.oov_token] * self.num_oov_indices
        return head + list(self.vocabulary)

    def vocabulary_size(self) -> int:
        return self._offset + len(self.vocabulary)

    vocab_size = vocabulary_size  # TF 2.3 / 2.4 name (spec.py:163-166)

    def _one(self, value) -> int:
        value = self._norm(value)
        if self.mask_token is not None and value == self.mask_token:
            return 0
        if value in self._index:
            return self._index[value]
        if self.num_oov_indices == 0:
            raise io_lib.InvalidArgumentError(io_lib.ERR_OOV, "value %r is not in the lookup vocabulary and num_oov_indices is 0" % (value,))
        return 0 if self.mask_token is None else 1

    def __call__(self, inputs):
        if isinstance(inputs, (str, bytes, int, np.integer)):
            return np.int64(self._one(inputs))
        arr = np.asarray(inputs, dtype=object)
        out = np.empty(arr.shape, dtype=np.int64)
        for idx in np.ndindex(arr.shape):
            out[idx] = self._one(arr[idx])
        return out


class StringLookup(_Lookup):
    oov_token = "[UNK]"

    @staticmethod
    def _norm(v):
        return v.decode("utf-8") if isinstance(v, bytes) else str(v)


class IntegerLookup(_Lookup):
    oov_token = -1

    @staticmethod
    def _norm(v):
        return int(v)


class SequenceDiscretizer:
    """``SequenceDiscretizer`` (discretizer.py:6-31): float32 cast, then Bucketize = number of float32 boundaries <= x."""

    def __init__(self, bins: Sequence[float]):
        self.bin_boundaries = [float(b) for b in bins]
        self._f32 = np.asarray(self.bin_boundaries, dtype=np.float32)

    def __call__(self, inputs):
        x = np.asarray(inputs).astype(np.float32)
        return np.searchsorted(self._f32, x, side="right").astype(np.int64)


# ----------------------------------------------------------------------------------------------------------------------------------
# tf.train.SequenceExample encoder (used to export datasets; the tests cross-check it against google.protobuf).
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_feature(values, dtype: str) -> bytes:
    """One ``tf.train.Feature`` of the kind the column dtype selects."""
    kind = _DTYPES[dtype]
    if kind == io_lib.STRING:
        body = b"".join(_ld(1, v if isinstance(v, bytes) else str(v).encode("utf-8")) for v in values)
        return _ld(1, body)
    if kind == io_lib.FLOAT32:
        packed = np.asarray(values, dtype="<f4").tobytes()
        return _ld(2, _ld(1, packed) if packed else b"")
    packed = b"".join(_varint(int(v)) for v in values)
    return _ld(3, _ld(1, packed) if packed else b"")


def encode_sequence_example(context: Dict[str, bytes], feature_lists: Dict[str, List[bytes]]) -> bytes:
    """``context``: key -> encoded Feature; ``feature_lists``: key -> encoded Feature per step."""
    ctx = b"".join(_ld(1, _ld(1, k.encode("utf-8")) + _ld(2, f)) for k, f in context.items())
    fls = b"".join(_ld(1, _ld(1, k.encode("utf-8")) + _ld(2, b"".join(_ld(1, f) for f in steps))) for k, steps in feature_lists.items())
    return _ld(1, ctx) + _ld(2, fls)


def write_tfrecord(path: str, records: Sequence[bytes]):
    lib = io_lib.load_library()
    n = len(records)
    ptrs = (ctypes.c_char_p * max(n, 1))(*records)
    lens = (ctypes.c_uint64 * max(n, 1))(*[len(r) for r in records])
    io_lib.check(lib.fdio_tfrecord_write(path.encode(), ptrs, lens, n))


def set_visual_default(decoded_data: Dict) -> Dict:
    """``data/spec.py:16-21`` (demo_crello's "attr" view): every element of an unbatched document drawn black, opaque, in a dummy font."""
    for element in decoded_data["elements"]:
        element["color"] = [0.0, 0.0, 0.0]
        element["opacity"] = 1.0
        element["font_family"] = "DummyFont"
    return decoded_data


class TFRecordFile:
    """One mmapped shard with its record index (``tf.data.TFRecordDataset`` over one file)."""

    def __init__(self, path: str, verify_crc: int = 2):
        self._lib = io_lib.load_library()
        self._handle = io_lib.check_handle(self._lib.fdio_tfrecord_open(path.encode(), int(verify_crc)))
        self.path = path
        n = self._lib.fdio_tfrecord_count(self._handle)
        self.pointers = np.empty(n, dtype=np.uint64)
        self.lengths = np.empty(n, dtype=np.uint64)
        p, ln = ctypes.c_void_p(), ctypes.c_uint64()
        for i in range(n):
            io_lib.check(self._lib.fdio_tfrecord_get(self._handle, i, ctypes.byref(p), ctypes.byref(ln)))
            self.pointers[i] = p.value or 0
            self.lengths[i] = ln.value

    def __len__(self):
        return len(self.pointers)

    def record(self, i: int) -> bytes:
        return ctypes.string_at(int(self.pointers[i]), int(self.lengths[i]))

    def close(self):
        if self._handle:
            self._lib.fdio_tfrecord_close(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------------------------------------------
class DataSpec:
    """Utility class to handle the data schema (spec.py:25-362)."""

    def __init__(self, name: str, path: str, batch_size: int = 8, num_threads: Optional[int] = None):
        self._path = path
        self._batch_size = batch_size
        self._threads = int(num_threads or min(16, os.cpu_count() or 1))
        if name in BUILTIN_SPECS:
            self._spec = BUILTIN_SPECS[name]
        elif os.path.exists(name):
            import yaml

            with open(name) as f:
                self._spec = yaml.safe_load(f)
        else:
            raise FileNotFoundError("No spec named %r (known: %s) and no such YAML file" % (name, ", ".join(BUILTIN_SPECS)))
        self._splits = self._load_resource("count.json")
        self._init_preprocessor()
        self._lib = io_lib.load_library()
        self._keepalive = []
        self._order = list(self.columns.keys())
        self._schemas = {}  # strings flag -> (native schema, output kind per column)

    # -------------------------------------------------------------------------------------------------------- schema / preprocessors
    @property
    def columns(self) -> Dict:
        return self._spec.get("columns", {})

    @property
    def preprocessor(self) -> Dict:
        return self._preprocessor

    def _init_preprocessor(self):
        """spec.py:90-105."""
        vocabulary = self._load_resource("vocabulary.json")
        self._preprocessor = OrderedDict()
        for name, column in self.columns.items():
            if "lookup" in column:
                self._preprocessor[name] = self._create_lookup(name, column, vocabulary)
            elif "discretize" in column:
                spec = column["discretize"]
                boundaries = list(np.linspace(spec["min"], spec["max"], spec["bins"]))[1:]
                self._preprocessor[name] = SequenceDiscretizer(boundaries)

    def _create_lookup(self, name, column, vocabulary):
        """spec.py:107-134: vocabulary.json entry (list, or token -> frequency filtered by min_freq) or an integer [min, max] range."""
        lookup = column["lookup"]
        assert name in vocabulary or (isinstance(lookup, dict) and "vocabulary" in lookup), name
        layer_fn = {"string": StringLookup, "int64": IntegerLookup}[column["dtype"]]
        if name in vocabulary:
            vocab = vocabulary[name]
        else:
            vocab = list(range(lookup["vocabulary"]["min"], lookup["vocabulary"]["max"] + 1))
        if isinstance(vocab, dict):
            vocab = [int(key) if column["dtype"] == "int64" else key for key, value in vocab.items() if value >= column.get("min_freq", 1)]
        options = {} if lookup is True else {k: v for k, v in lookup.items() if k != "vocabulary"}
        return layer_fn(vocabulary=vocab, **options)

    def _schema(self, strings: bool):
        """The native schema; ``strings=False`` validates the raw byte-string columns (``id``, ``uuid``) without emitting them."""
        if strings in self._schemas:
            return self._schemas[strings]
        cols = (io_lib.Column * len(self.columns))()
        keep = self._keepalive
        out_kind = {}
        for i, (name, column) in enumerate(self.columns.items()):
            c = cols[i]
            kind = _DTYPES[column["dtype"]]
            c.name = name.encode("utf-8")
            c.is_sequence = 1 if column.get("is_sequence") else 0
            c.dtype = kind
            c.width = int(np.prod(column.get("shape", (1,))))
            layer = self._preprocessor.get(name)
            if isinstance(layer, _Lookup):
                c.transform = io_lib.LOOKUP
                c.vocab_size = len(layer.vocabulary)
                c.num_oov_indices = layer.num_oov_indices
                c.has_mask = 0 if layer.mask_token is None else 1
                if kind == io_lib.STRING:
                    arr = (ctypes.c_char_p * max(1, c.vocab_size))(*[v.encode("utf-8") for v in layer.vocabulary])
                    c.vocab_str = arr
                    if layer.mask_token is not None:
                        c.mask_str = layer.mask_token.encode("utf-8")
                else:
                    arr = (ctypes.c_int64 * max(1, c.vocab_size))(*layer.vocabulary)
                    c.vocab_int = arr
                    c.mask_int = layer.mask_token or 0
                keep.append(arr)
                c.output = io_lib.OUT_INT32
            elif isinstance(layer, SequenceDiscretizer):
                c.transform = io_lib.DISCRETIZE
                arr = (ctypes.c_float * len(layer.bin_boundaries))(*layer.bin_boundaries)
                keep.append(arr)
                c.n_boundaries = len(layer.bin_boundaries)
                c.boundaries = arr
                c.output = io_lib.OUT_INT32
            else:
                c.transform = io_lib.NONE
                c.output = {io_lib.INT64: io_lib.OUT_INT32, io_lib.FLOAT32: io_lib.OUT_FLOAT32,
                            io_lib.STRING: io_lib.OUT_SPAN if strings else io_lib.OUT_SKIP}[kind]
            out_kind[name] = c.output
        handle = io_lib.check_handle(self._lib.fdio_schema_create(cols, len(self.columns)))
        self._schemas[strings] = (handle, out_kind)
        return self._schemas[strings]

    def __del__(self):
        try:
            for handle, _ in self._schemas.values():
                self._lib.fdio_schema_destroy(handle)
            self._schemas = {}
        except Exception:
            pass

    def size(self, split: str) -> int:
        """Length of the records for the split."""
        return self._splits[split]

    def steps_per_epoch(self, split: str, batch_size: Optional[int] = None) -> int:
        return int(np.ceil(self.size(split) / (batch_size or self._batch_size)))

    def make_input_columns(self) -> Dict:
        """Input specification for a model (spec.py:144-211)."""
        inputs = OrderedDict()
        for key, column in self.columns.items():
            layer = self._preprocessor.get(key)
            if column.get("demo_only", False):
                inputs[key] = {"demo_only": True}
            elif isinstance(layer, SequenceDiscretizer):
                inputs[key] = {"type": "categorical", "input_dim": len(layer.bin_boundaries) + 1}
            elif isinstance(layer, _Lookup):
                inputs[key] = {"type": "categorical", "input_dim": layer.vocabulary_size()}
            elif column["dtype"] in ("int", "int32", "int64"):
                inputs[key] = {"type": "categorical", "input_dim": column["max"] + 1}
            elif column["dtype"] in ("float", "float32", "float64"):
                inputs[key] = {"type": "numerical"}
            else:
                raise NotImplementedError
            inputs[key]["shape"] = tuple(column.get("shape", (1,)))
            inputs[key]["is_sequence"] = column.get("is_sequence", False)
            if "primary_label" in column:
                inputs[key]["primary_label"] = int(self._preprocessor[key](column["primary_label"]["default"]))
            else:
                inputs[key]["primary_label"] = None
        for key, column in self.columns.items():
            if "loss_condition" in column:
                cond = column["loss_condition"]
                mask = [v in cond["values"] for v in self._preprocessor[cond["key"]].get_vocabulary()]
                inputs[key]["loss_condition"] = {"key": cond["key"], "mask": mask}
        return inputs

    # ---------------------------------------------------------------------------------------------------------------- parsing
    def parse_records(self, pointers: np.ndarray, lengths: np.ndarray, pad_to: Optional[int] = None, pin_memory: bool = False,
                      strings: bool = True, packed: bool = False) -> Dict:
        """Parses ``B`` serialized SequenceExamples given by address and length (spec.py:255-287).  Sequence columns come out
        ``(B, S, *shape)`` with ``S`` = the longest document of the batch (``parse_sequence_example`` semantics) or ``pad_to``.

        ``packed=True`` writes the numerical sequence columns in the input pipeline's packed format straight from the records
        (``fdio_parse_batch_packed``): ``batch[key]`` is ``float32 [n_rows, C]`` -- only the rows of the elements that carry the field --
        and ``batch[key + "/rows"]`` the ``int32 [B, S]`` element -> row map; bit-identical to ``flex_dm_b200.data.pack_batch`` of the
        dense batch, without ever writing (or copying to the device) the 59 % of crello's embedding rows nothing reads."""
        B = len(pointers)
        schema, out_kind = self._schema(bool(strings))
        ptrs, lens, _keep = self._record_arrays(pointers, lengths)
        if pad_to is None:  # parse_sequence_example pads to the longest document of the batch: one cheap pass over the record structure
            steps = (ctypes.c_int32 * max(B, 1))()
            io_lib.check(self._lib.fdio_batch_steps(schema, ptrs, lens, B, steps, self._threads))
            S = max(steps[:B]) if B else 0
        else:  # fixed shape: the fill pass itself reports a document that does not fit
            S = int(pad_to)
        out_ptrs = (ctypes.c_void_p * len(self._order))()
        output, spans = OrderedDict(), {}
        pack_keys = self._packed_columns() if packed else {}
        capacity, rowmaps = {}, {}
        for i, name in enumerate(self._order):
            column = self.columns[name]
            shape = tuple(column.get("shape", (1,)))
            full = (B, S) + shape if column.get("is_sequence") else (B,) + shape
            kind = out_kind[name]
            if kind == io_lib.OUT_SKIP:
                continue
            if name in pack_keys:  # rows in use only: capacity = every element (pinned pages nothing writes are never touched)
                t = capacity[name] = torch.empty((max(B * S, 1), int(np.prod(shape))), dtype=torch.float32, pin_memory=pin_memory)
                rowmaps[name] = torch.empty((B, S), dtype=torch.int32, pin_memory=pin_memory)
                output[name] = t
                out_ptrs[i] = t.data_ptr()
                continue
            if kind == io_lib.OUT_SPAN:
                t = spans[name] = torch.zeros(full + (2,), dtype=torch.int64)
            else:
                t = torch.empty(full, dtype=torch.int32 if kind == io_lib.OUT_INT32 else torch.float32, pin_memory=pin_memory)  # every slot is written
            output[name] = t
            out_ptrs[i] = t.data_ptr()
        if pack_keys:
            from .data import ROWS_SUFFIX

            order = list(self._order)
            pack = (io_lib.PackColumn * len(pack_keys))()
            keep = []
            for k, (name, cond) in enumerate(pack_keys.items()):
                pack[k].column = order.index(name)
                pack[k].cond_column = order.index(cond["key"]) if cond else -1
                if cond:
                    mask = np.ascontiguousarray(np.asarray(cond["mask"], dtype=np.uint8))
                    keep.append(mask)
                    pack[k].cond_mask = mask.ctypes.data_as(io_lib.c_u8p)
                    pack[k].cond_n = int(mask.size)
                pack[k].rowmap = ctypes.cast(rowmaps[name].data_ptr(), ctypes.POINTER(ctypes.c_int32))
                pack[k].capacity_rows = int(capacity[name].shape[0])
            code = self._lib.fdio_parse_batch_packed(schema, ptrs, lens, B, S, out_ptrs, order.index("length"), pack, len(pack_keys), self._threads)
            if code == io_lib.OK:
                for k, name in enumerate(pack_keys):
                    output[name] = capacity[name][: int(pack[k].n_rows)]
                    output[name + ROWS_SUFFIX] = rowmaps[name]
        else:
            code = self._lib.fdio_parse_batch(schema, ptrs, lens, B, S, out_ptrs, self._threads)
        if code == io_lib.ERR_ARG and pad_to is not None and "more steps" in io_lib.last_error():
            raise ValueError("A document has more than pad_to=%d elements (%s)" % (pad_to, io_lib.last_error()))
        io_lib.check(code)
        for name, t in spans.items():  # (offset, length) pairs -> byte strings
            sp = t.numpy()
            arr = np.empty(sp.shape[:-1], dtype=object)
            for idx in np.ndindex(arr.shape):
                off, n = int(sp[idx][0]), int(sp[idx][1])
                arr[idx] = ctypes.string_at(int(pointers[idx[0]]) + off, n) if n else b""
            output[name] = arr
        return output

    def _packed_columns(self) -> Dict:
        """Numerical sequence columns the packed batch format applies to -> their loss_condition ({"key", "mask"}) or None
        (the same selection as ``flex_dm_b200.data.pack_batch``)."""
        if getattr(self, "_pack_keys", None) is None:
            keys = OrderedDict()
            for key, column in self.make_input_columns().items():
                if column.get("is_sequence") and column.get("type") == "numerical" and not column.get("demo_only", False) and key in self._order:
                    keys[key] = column.get("loss_condition")
            self._pack_keys = keys
        return self._pack_keys

    @staticmethod
    def _record_arrays(pointers: np.ndarray, lengths: np.ndarray):
        """The native call's (record pointers, record lengths) arguments straight from the numpy index arrays (no per-record Python)."""
        p = np.ascontiguousarray(pointers, dtype=np.uint64)
        n = np.ascontiguousarray(lengths, dtype=np.uint64)
        if p.size == 0:
            p, n = np.zeros(1, np.uint64), np.zeros(1, np.uint64)
        return p.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p)), n.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), (p, n)

    def record_steps(self, pointers: np.ndarray, lengths: np.ndarray) -> np.ndarray:
        """Number of elements (sequence steps) of each record."""
        B = len(pointers)
        schema, _ = self._schema(False)
        ptrs, lens, _keep = self._record_arrays(pointers, lengths)
        steps = (ctypes.c_int32 * max(B, 1))()
        io_lib.check(self._lib.fdio_batch_steps(schema, ptrs, lens, B, steps, self._threads))
        return np.asarray(steps[:B], dtype=np.int64)

    def pad_word(self, name: str) -> int:
        """The 32-bit pattern a padded step of a sequence column holds: the parse default (0 / 0.0 / "") through its preprocessor."""
        column = self.columns[name]
        layer = self._preprocessor.get(name)
        if layer is None:
            return 0
        default = "" if column["dtype"] == "string" else 0
        return int(np.asarray(layer([default] if isinstance(layer, SequenceDiscretizer) else default)).reshape(-1)[0])

    def parse_fn(self, serialized: Sequence[bytes], pad_to: Optional[int] = None) -> Dict:
        """``DataSpec.parse_fn`` over a batch of serialized SequenceExample byte strings (spec.py:255-287)."""
        bufs = [ctypes.create_string_buffer(s, len(s)) for s in serialized]
        pointers = np.asarray([ctypes.addressof(b) for b in bufs], dtype=np.uint64)
        lengths = np.asarray([len(s) for s in serialized], dtype=np.uint64)
        return self.parse_records(pointers, lengths, pad_to=pad_to)

    # ---------------------------------------------------------------------------------------------------------------- datasets
    def make_dataset(self, split: str, batch_size: Optional[int] = None, shuffle=None, repeat: bool = False, prefetch: Optional[int] = 2,
                     parallel=None, cache=None, seed: int = 0, pad_to: Optional[int] = None, pin_memory: Optional[bool] = None,
                     strings: bool = False, verify_crc: int = 1, device=None, shard=None, packed: bool = False) -> "RecordDataset":
        """spec.py:213-253: list ``<split>-*.tfrecord``, read, [shuffle], [repeat], batch, parse, prefetch.

        ``parallel`` and ``cache=True`` are accepted for signature compatibility: shards are always mmapped (the page cache is the cache)
        and parsing always uses the spec's host threads.  ``cache="device"`` parses the split once into HBM and cuts every batch out of
        it on the GPU (``device_cache.DeviceCachedDataset``; same batches for the same ``seed``).  ``shard=(rank, world_size)`` is the
        document-sharded data-parallel split (``tf.data``'s ``shard``): every rank draws the same (seeded) document
        order and keeps every ``world_size``-th document starting at ``rank``, so ranks see disjoint documents and equally many batches.  ``shuffle=True`` shuffles over the whole split like the reference
        (``shuffle = self.size(split)``); an integer is a shuffle-buffer size.  ``strings=False`` leaves the demo-only byte-string
        columns (``id``, ``uuid``) out of the batches -- ``MFP`` drops them anyway (mfp.py:235-237).  ``packed=True`` yields batches in the packed column format
        (``parse_records``): what ``DevicePrefetcher`` / ``MFP.train_step`` take directly, 55 MB instead of 135 MB per 256 x 128 crello batch."""
        assert split in self._splits, "split must be one of (%s)" % ", ".join(self._splits.keys())
        if shuffle is True:
            shuffle = self.size(split)
        pattern = os.path.join(self._path, split + "-*.tfrecord")
        files = sorted(glob.glob(pattern))
        if not files:
            raise FileNotFoundError("No TFRecord matches %s" % pattern)
        if pin_memory is None:
            pin_memory = torch.cuda.is_available()
        if cache == "device":  # the parsed split resident in HBM, batches gathered on the GPU (device_cache.py)
            from .device_cache import DeviceCachedDataset

            return DeviceCachedDataset(RecordDataset(self, files, batch_size or self._batch_size, int(shuffle or 0), repeat, 0, seed, pad_to, pin_memory,
                                                     False, verify_crc, shard), device=device)
        return RecordDataset(self, files, batch_size or self._batch_size, int(shuffle or 0), repeat, prefetch or 0, seed, pad_to, pin_memory,
                             strings, verify_crc, shard, packed=packed)

    # ---------------------------------------------------------------------------------------------------------------- post-processing
    def logit_to_label(self, example: Dict) -> Dict:
        """Convert logit prediction to labels (spec.py:289-299)."""
        for key, column in self.columns.items():
            if column.get("demo_only", False) or key not in example:
                continue
            rank = 1 + int(bool(column.get("is_sequence", 0))) + len(column.get("shape", (1,)))
            x = example[key]
            if x.ndim >= rank + 1:
                x = torch.as_tensor(x)
                example[key] = torch.argmax(x, dim=-1).to(torch.int32)
        return example

    def unbatch(self, example: Dict) -> List[Dict]:
        """A batch -> list of items ``{key: value, "elements": [{key: value}]}`` with lookups / discretisation undone (spec.py:301-346)."""
        example = self.logit_to_label(dict(example))
        host = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in example.items()}
        batch_size = host["length"].shape[0]
        items = []
        for i in range(batch_size):
            length = int(np.squeeze(host["length"][i]) + 1)  # zero-based
            for name, column in self.columns.items():
                if column.get("is_sequence") and name in host:
                    length = min(length, host[name][i].shape[0])
                    break
            item = {"elements": [{} for _ in range(length)]}
            for name, column in self.columns.items():
                if name not in host:
                    continue
                x = host[name][i]
                if "lookup" in column:
                    table = np.array(self._preprocessor[name].get_vocabulary())
                    x = table[x]
                elif "discretize" in column:
                    spec = column["discretize"]
                    scale = (spec["max"] - spec["min"]) / (spec["bins"] - 1.0)
                    x = scale * x + spec["min"]
                if column.get("is_sequence"):
                    for j in range(length):
                        item["elements"][j][name] = x[j, :].tolist() if x.shape[1] > 1 else x[j, 0]
                else:
                    item[name] = x[0]
            items.append(item)
        return items

    def _load_resource(self, path: str):
        with open(os.path.join(self._path, path)) as f:
            return json.load(f)


class RecordDataset:
    """The iterable ``make_dataset`` returns: an index stream over the mmapped shards (shuffle / repeat), cut into batches, parsed by the
    native library on a producer thread (``prefetch`` batches ahead; the ctypes call releases the GIL) into fresh pinned tensors from
    torch's caching host allocator, which recycles them safely under asynchronous device copies."""

    def __init__(self, spec: DataSpec, files: List[str], batch_size: int, shuffle: int, repeat: bool, prefetch: int, seed: int,
                 pad_to: Optional[int], pin_memory: bool, strings: bool, verify_crc: int, shard=None, packed: bool = False):
        self.spec = spec
        self.shards = [TFRecordFile(f, verify_crc) for f in files]
        self.pointers = np.concatenate([s.pointers for s in self.shards]) if self.shards else np.empty(0, np.uint64)
        self.lengths = np.concatenate([s.lengths for s in self.shards]) if self.shards else np.empty(0, np.uint64)
        self.batch_size, self.shuffle, self.repeat, self.prefetch = int(batch_size), int(shuffle), bool(repeat), int(prefetch)
        self.seed, self.pad_to, self.pin_memory, self.strings = seed, pad_to, pin_memory, strings
        self.packed = bool(packed)
        self._epoch = 0
        self.shard = None
        if shard is not None:
            rank, world = int(shard[0]), int(shard[1])
            if not 0 <= rank < world:
                raise ValueError("shard=(rank, world_size) with 0 <= rank < world_size, got %r" % (shard,))
            self.shard = (rank, world)

    def __len__(self):
        return len(self.pointers)

    def _index_stream(self, rng: np.random.Generator) -> Iterator[int]:
        if self.shard is None:
            yield from self._global_stream(rng)
            return
        rank, world = self.shard
        if self.repeat:
            for k, i in enumerate(self._global_stream(rng)):  # identical on every rank (same seed): keep every world-th document
                if k % world == rank:
                    yield i
            return
        # one finite pass (val / test): every rank gets ceil(n / world) documents -- the last round wraps to the head of the pass -- so that all
        # ranks run the same number of batches (metric rows are all-reduced collectively; an uneven split would hang or lose a batch)
        order = list(self._global_stream(rng))
        per = -(-len(order) // world)
        for j in range(per):
            yield order[(j * world + rank) % len(order)]

    def _global_stream(self, rng: np.random.Generator) -> Iterator[int]:
        n = len(self.pointers)
        while True:
            order = np.arange(n)
            if self.shuffle >= n:
                rng.shuffle(order)  # a buffer as large as the split (shuffle=True) is a full permutation per pass
                yield from order.tolist()
            elif self.shuffle > 1:  # tf.data shuffle-buffer semantics
                buf = order[: self.shuffle].tolist()
                for nxt in order[self.shuffle:].tolist():
                    k = int(rng.integers(len(buf)))
                    yield buf[k]
                    buf[k] = nxt
                while buf:
                    yield buf.pop(int(rng.integers(len(buf))))
            else:
                yield from order.tolist()
            if not self.repeat:
                return

    def _batches(self) -> Iterator[Dict]:
        rng = np.random.Generator(np.random.PCG64([self.seed, self._epoch]))
        self._epoch += 1
        picked: List[int] = []
        for i in self._index_stream(rng):
            picked.append(i)
            if len(picked) == self.batch_size:
                yield self._parse(picked)
                picked = []
        if picked:
            yield self._parse(picked)

    def _parse(self, picked: List[int]) -> Dict:
        idx = np.asarray(picked, dtype=np.int64)
        return self.spec.parse_records(self.pointers[idx], self.lengths[idx], pad_to=self.pad_to, pin_memory=self.pin_memory, strings=self.strings,
                                       packed=self.packed)

    def __iter__(self) -> "BatchIterator":
        if self.prefetch <= 0:
            return BatchIterator(self._batches())
        return BatchIterator(_Prefetched(self._batches(), self.prefetch))


class BatchIterator:
    """``iter(dataset)``: a Python iterator that also answers ``tf.data``'s ``iterator.get_next()`` (eval.py:47-49).  Where TensorFlow
    raises ``OutOfRangeError`` at the end of a non-repeating dataset this raises ``StopIteration``."""

    def __init__(self, source: Iterator[Dict]):
        self._source = source

    def __iter__(self) -> "BatchIterator":
        return self

    def __next__(self) -> Dict:
        return next(self._source)

    get_next = __next__

    def close(self) -> None:
        close = getattr(self._source, "close", None)
        if close is not None:
            close()


class _Prefetched:
    """``dataset.prefetch(n)``: a producer thread keeps ``n`` parsed batches ahead of the consumer."""

    _END = object()

    def __init__(self, source: Iterable, depth: int):
        self._q: "queue.Queue" = queue.Queue(maxsize=depth)
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, args=(iter(source),), daemon=True)
        self._thread.start()

    def _put(self, item) -> bool:
        while not self._stop.is_set():
            try:
                self._q.put(item, timeout=0.1)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, it):
        try:
            for item in it:
                if not self._put(item):
                    return
            self._put(self._END)
        except BaseException as e:  # surfaced in the consumer
            self._put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            self._q.put(self._END)
            raise StopIteration
        if isinstance(item, BaseException):
            self._q.put(self._END)
            raise item
        return item

    def close(self):
        self._stop.set()

    def __del__(self):
        self._stop.set()
