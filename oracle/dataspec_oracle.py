"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input side, used as the checker of ``libflexdm_io.so``.

Follows ``src/mfp/mfp/data/spec.py``: ``_init_preprocessor`` (:90-105), ``_create_lookup`` (:107-134), ``parse_fn`` (:255-287) and
``data/discretizer.py:6-31``, with the TensorFlow pieces underneath restated from their published formats / documented behaviour
(TensorFlow is not under /root/reference; **parity unpinned** for these primitives):

* TFRecord framing (``tensorflow/core/lib/io/record_writer.h``) and CRC-32C (RFC 3720 test vectors are checked in the tests);
* ``tf.train.SequenceExample`` wire format (``tensorflow/core/example/*.proto``) -- decoded here with a separate pure-Python reader;
* ``tf.io.parse_sequence_example`` with ``FixedLenFeature`` / ``FixedLenSequenceFeature``: pad to the batch maximum with 0 / 0.0 / b"";
* Keras ``StringLookup`` / ``IntegerLookup`` (``[mask] + [OOV] + vocabulary``) and ``Discretization`` (Bucketize on float32).

Pure Python + numpy, written independently of the C++ (bitwise CRC, recursive-descent wire reader); sized for test batches only.
"""
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np


# ---- CRC-32C, bit by bit ---------------------------------------------------------------------------------------------------------
def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    for byte in data:
        c ^= byte
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- TFRecord framing ------------------------------------------------------------------------------------------------------------
def read_tfrecord(path: str, verify: bool = True) -> List[bytes]:
    out = []
    with open(path, "rb") as f:
        blob = f.read()
    pos = 0
    while pos < len(blob):
        (n,) = struct.unpack_from("<Q", blob, pos)
        (len_crc,) = struct.unpack_from("<I", blob, pos + 8)
        data = blob[pos + 12:pos + 12 + n]
        (data_crc,) = struct.unpack_from("<I", blob, pos + 12 + n)
        if verify:
            assert mask_crc(crc32c(blob[pos:pos + 8])) == len_crc, "length crc"
            assert mask_crc(crc32c(data)) == data_crc, "data crc"
        out.append(data)
        pos += 16 + n
    return out


def frame_record(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data)))


# ---- protobuf wire reader --------------------------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = value = 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7


def _fields(buf: bytes) -> List[Tuple[int, int, object]]:
    """[(field, wire type, value)] with value = int (varint / fixed) or bytes (length-delimited)."""
    out, pos = [], 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("wire type %d" % wt)
        out.append((field, wt, v))
    return out


def decode_feature(buf: bytes):
    """-> ("string" | "float32" | "int64" | None, list of values)."""
    kind, values = None, []
    for field, wt, v in _fields(buf):
        if wt != 2 or field not in (1, 2, 3):
            continue
        kind, values = {1: "string", 2: "float32", 3: "int64"}[field], []
        for f2, t2, v2 in _fields(v):
            if f2 != 1:
                continue
            if kind == "string":
                values.append(v2)
            elif kind == "float32":
                values.extend(np.frombuffer(v2, dtype="<f4").tolist() if t2 == 2 else [struct.unpack("<f", v2)[0]])
            else:
                if t2 == 2:
                    pos = 0
                    while pos < len(v2):
                        u, pos = _varint(v2, pos)
                        values.append(u - (1 << 64) if u >= (1 << 63) else u)
                else:
                    values.append(v2 - (1 << 64) if v2 >= (1 << 63) else v2)
    return kind, values


def _decode_map(buf: bytes) -> Dict[str, bytes]:
    out = {}
    for field, wt, entry in _fields(buf):
        if field != 1 or wt != 2:
            continue
        key, val = "", b""
        for f2, _, v2 in _fields(entry):
            if f2 == 1:
                key = v2.decode("utf-8")
            elif f2 == 2:
                val = v2
        out[key] = val
    return out


def decode_sequence_example(buf: bytes):
    """-> (context: key -> (kind, values), feature_lists: key -> [(kind, values) per step])."""
    context, lists = {}, {}
    for field, wt, v in _fields(buf):
        if field == 1 and wt == 2:
            context.update({k: decode_feature(f) for k, f in _decode_map(v).items()})
        elif field == 2 and wt == 2:
            for k, fl in _decode_map(v).items():
                lists[k] = [decode_feature(f) for f1, t1, f in _fields(fl) if f1 == 1 and t1 == 2]
    return context, lists


# ---- preprocessors -----------------------------------------------------------------------------------------------------------------
class Lookup:
    """spec.py:107-134 + Keras index layout: [mask_token] + [OOV] * num_oov_indices + vocabulary."""

    def __init__(self, column: Dict, name: str, vocabulary: Dict):
        lookup = column["lookup"]
        if name in vocabulary:
            vocab = vocabulary[name]
        else:
            vocab = list(range(lookup["vocabulary"]["min"], lookup["vocabulary"]["max"] + 1))
        if isinstance(vocab, dict):
            vocab = [int(k) if column["dtype"] == "int64" else k for k, v in vocab.items() if v >= column.get("min_freq", 1)]
        options = {} if lookup is True else {k: v for k, v in lookup.items() if k != "vocabulary"}
        self.num_oov = options.get("num_oov_indices", 1)
        self.mask = options.get("mask_token", options.get("mask_value", None))
        self.is_string = column["dtype"] == "string"
        self.tokens = ([] if self.mask is None else [self.mask]) + [("[UNK]" if self.is_string else -1)] * self.num_oov + list(vocab)

    def __call__(self, value):
        if self.is_string and isinstance(value, bytes):
            value = value.decode("utf-8")
        if self.mask is not None and value == self.mask:
            return 0
        first = (0 if self.mask is None else 1) + self.num_oov
        for i in range(first, len(self.tokens)):
            if self.tokens[i] == value:
                return i
        if self.num_oov == 0:
            raise KeyError(value)
        return first - 1


class Discretizer:
    """spec.py:97-105 + discretizer.py:20-25: boundaries = linspace(min, max, bins)[1:]; float32 cast; Bucketize."""

    def __init__(self, column: Dict):
        spec = column["discretize"]
        self.boundaries = [np.float32(b) for b in list(np.linspace(spec["min"], spec["max"], spec["bins"]))[1:]]

    def __call__(self, value):
        x = np.float32(value)
        return sum(1 for b in self.boundaries if b <= x)


def make_preprocessors(columns: Dict, vocabulary: Dict) -> Dict:
    pre = {}
    for name, column in columns.items():
        if "lookup" in column:
            pre[name] = Lookup(column, name, vocabulary)
        elif "discretize" in column:
            pre[name] = Discretizer(column)
    return pre


# ---- parse_fn ----------------------------------------------------------------------------------------------------------------------
def parse_fn(columns: Dict, vocabulary: Dict, serialized: Sequence[bytes]) -> Dict[str, np.ndarray]:
    """spec.py:255-287 over a list of serialized SequenceExamples."""
    pre = make_preprocessors(columns, vocabulary)
    decoded = [decode_sequence_example(s) for s in serialized]
    B = len(serialized)
    out = {}
    for name, column in columns.items():
        width = int(np.prod(column.get("shape", (1,))))
        shape = tuple(column.get("shape", (1,)))
        dtype = {"int64": "int64", "float32": "float32", "string": "string"}[column["dtype"]]
        default = {"int64": 0, "float32": 0.0, "string": b""}[dtype]
        fn = pre.get(name, lambda v: v)
        if column.get("is_sequence"):
            steps = []
            for _, lists in decoded:
                assert name in lists, "feature list %s is required" % name
                rows = []
                for kind, values in lists[name]:
                    assert kind == dtype and len(values) == width, (name, kind, len(values))
                    rows.append(values)
                steps.append(rows)
            S = max((max(len(fl) for fl in lists.values()) if lists else 0) for _, lists in decoded) if B else 0
            raw = [[rows[t] if t < len(rows) else [default] * width for t in range(S)] for rows in steps]
            full = (B, S) + shape
        else:
            raw = []
            for context, _ in decoded:
                assert name in context, "feature %s is required" % name
                kind, values = context[name]
                assert kind == dtype and len(values) == width, (name, kind, len(values))
                raw.append(values)
            full = (B,) + shape
        flat = [fn(v) for v in np.asarray(raw, dtype=object).reshape(-1).tolist()] if B else []
        if name in pre or dtype == "int64":
            out[name] = np.asarray(flat, dtype=np.int64).astype(np.int32).reshape(full)  # int64 -> int32, spec.py:281-285
        elif dtype == "float32":
            out[name] = np.asarray(flat, dtype=np.float32).reshape(full)
        else:
            arr = np.empty(len(flat), dtype=object)
            arr[:] = flat
            out[name] = arr.reshape(full)
    return out
