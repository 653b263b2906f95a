"""TensorFlow-format weight files for ``MFP.load_weights`` / ``save_weights`` (reference: ``train.py:67-69,94-97``,
``eval.py:169-172``, ``helpers/callbacks.py:49-56`` -- ``best.ckpt`` / ``final.ckpt`` are Keras object-based checkpoints in the
tensor-bundle format: ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``).

The bundle files are read and written by ``libflexdm_io.so`` (``csrc/io/bundle.cc``; ``include/flexdm_io.h``).  This module holds what
sits above the bytes:

* the ``TrackableObjectGraph`` stored under ``_CHECKPOINTABLE_OBJECT_GRAPH`` (``tensorflow/core/protobuf/trackable_object_graph.proto``):
  nodes with named child edges, variables as leaves whose attribute names the bundle key.  Restoring follows TensorFlow's object-based
  matching: every variable of this model is found by walking its attribute path (``model/encoder/input_layer/left/embeddings`` ...,
  SURVEY.md Appendix B) edge by edge from the root -- independent of how the keys are spelled;
* a key-based fallback for bundles without an object graph (``<attribute path>/.ATTRIBUTES/VARIABLE_VALUE``), tolerant of extra wrapper
  edges as long as the match is unique and the shape agrees;
* the writer: variables + object graph + the ``checkpoint`` state file Keras leaves next to the bundle.

TensorFlow is not available here, so files written by TF itself could not be tested: the formats follow the published
specifications and are pinned by round trips and by hand-assembled fixtures in ``tests/test_io_formats.py``; the object-graph message and
the dtype numbers are also checked against the TensorFlow-authored protobuf classes inside the ``tensorboard`` package
(``tests/test_tf_authored_pins.py``).  The bundle's table format and ``BundleEntryProto`` stay "parity unpinned" (DESIGN.md section 2).
"""
import ctypes
import os
from collections import OrderedDict, deque
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import io_lib

OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VARIABLE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_, 17: np.uint16,
           19: np.float16, 22: np.uint32, 23: np.uint64}
DT_STRING = 7
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------------------------- tiny protobuf helpers
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = value = 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if not b & 0x80:
            return value, pos
        shift += 7


def _fields(buf: bytes):
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield field, wt, v


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v)


# ---------------------------------------------------------------------------------------------------------------- object graph
class ObjectGraph:
    """``TrackableObjectGraph``: ``children[i]`` = ordered {local_name: node id}, ``attributes[i]`` = {attribute name: checkpoint key}."""

    def __init__(self):
        self.children: List["OrderedDict[str, int]"] = []
        self.attributes: List["OrderedDict[str, str]"] = []
        self.full_names: List[Dict[str, str]] = []

    @classmethod
    def parse(cls, blob: bytes) -> "ObjectGraph":
        g = cls()
        for field, wt, node in _fields(blob):
            if field != 1 or wt != 2:
                continue
            ch, at, fn = OrderedDict(), OrderedDict(), {}
            for f2, t2, v2 in _fields(node):
                if f2 == 1 and t2 == 2:  # ObjectReference
                    node_id, name = 0, ""
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            node_id = v3
                        elif f3 == 2:
                            name = v3.decode("utf-8")
                    ch[name] = node_id
                elif f2 == 2 and t2 == 2:  # SerializedTensor
                    name = full = key = ""
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            name = v3.decode("utf-8")
                        elif f3 == 2:
                            full = v3.decode("utf-8")
                        elif f3 == 3:
                            key = v3.decode("utf-8")
                    at[name] = key
                    fn[name] = full
            g.children.append(ch)
            g.attributes.append(at)
            g.full_names.append(fn)
        return g

    def serialize(self) -> bytes:
        out = b""
        for ch, at, fn in zip(self.children, self.attributes, self.full_names):
            node = b"".join(_ld(1, (_vi(1, nid) if nid else b"") + _ld(2, name.encode("utf-8"))) for name, nid in ch.items())
            node += b"".join(_ld(2, _ld(1, name.encode("utf-8")) + _ld(2, fn.get(name, "").encode("utf-8")) + _ld(3, key.encode("utf-8")))
                             for name, key in at.items())
            out += _ld(1, node)
        return out

    def walk(self, path: List[str]) -> Optional[int]:
        node = 0
        for edge in path:
            if node >= len(self.children) or edge not in self.children[node]:
                return None
            node = self.children[node][edge]
        return node

    def variable_key(self, path: List[str]) -> Optional[str]:
        node = self.walk(path)
        if node is None or node >= len(self.attributes):
            return None
        return self.attributes[node].get("VARIABLE_VALUE")

    def variables(self) -> "OrderedDict[str, str]":
        """Every variable reachable from the root: shortest attribute path (breadth first, edge order) -> checkpoint key."""
        out, seen, todo = OrderedDict(), {0}, deque([(0, [])])
        while todo:
            node, path = todo.popleft()
            if node < len(self.attributes) and "VARIABLE_VALUE" in self.attributes[node]:
                out["/".join(path)] = self.attributes[node]["VARIABLE_VALUE"]
            for name, nid in (self.children[node].items() if node < len(self.children) else ()):
                if nid not in seen:
                    seen.add(nid)
                    todo.append((nid, path + [name]))
        return out

    @classmethod
    def from_variable_paths(cls, names: List[str]) -> "ObjectGraph":
        """The graph ``save_weights`` writes: one node per path prefix, variables as leaves (key = path + VARIABLE_SUFFIX)."""
        g = cls()
        index = {(): 0}
        g.children.append(OrderedDict())
        g.attributes.append(OrderedDict())
        g.full_names.append({})
        for name in names:
            parts = tuple(name.split("/"))
            for depth in range(1, len(parts) + 1):
                prefix = parts[:depth]
                if prefix not in index:
                    index[prefix] = len(g.children)
                    g.children.append(OrderedDict())
                    g.attributes.append(OrderedDict())
                    g.full_names.append({})
                    g.children[index[prefix[:-1]]][prefix[-1]] = index[prefix]
            leaf = index[parts]
            g.attributes[leaf]["VARIABLE_VALUE"] = name + VARIABLE_SUFFIX
            g.full_names[leaf]["VARIABLE_VALUE"] = "/".join(parts[-2:])
        return g


# ---------------------------------------------------------------------------------------------------------------- bundle access
class Bundle:
    """Read access to one checkpoint prefix."""

    def __init__(self, prefix: str):
        self._lib = io_lib.load_library()
        self._handle = io_lib.check_handle(self._lib.fdio_bundle_open(prefix.encode()))
        self.prefix = prefix
        self.keys = [self._lib.fdio_bundle_key(self._handle, i).decode("utf-8") for i in range(self._lib.fdio_bundle_count(self._handle))]
        self._index = {k: i for i, k in enumerate(self.keys)}

    def close(self):
        if self._handle:
            self._lib.fdio_bundle_close(self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __contains__(self, key: str) -> bool:
        return key in self._index

    def info(self, key: str) -> Tuple[int, Tuple[int, ...], int]:
        i = self._index[key]
        dtype, rank, nbytes = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        dims = (ctypes.c_int64 * 8)()
        io_lib.check(self._lib.fdio_bundle_info(self._handle, i, ctypes.byref(dtype), ctypes.byref(rank), dims, 8, ctypes.byref(nbytes)))
        if rank.value > 8:
            raise io_lib.IOError_(io_lib.ERR_UNSUPPORTED, "tensor '%s' has rank %d" % (key, rank.value))
        return dtype.value, tuple(dims[k] for k in range(rank.value)), nbytes.value

    def raw(self, key: str) -> bytes:
        _, _, nbytes = self.info(key)
        buf = ctypes.create_string_buffer(max(nbytes, 1))
        io_lib.check(self._lib.fdio_bundle_read(self._handle, self._index[key], buf, nbytes))
        return buf.raw[:nbytes]

    def tensor(self, key: str) -> np.ndarray:
        dtype, shape, nbytes = self.info(key)
        if dtype not in _DTYPES:
            raise io_lib.IOError_(io_lib.ERR_UNSUPPORTED, "tensor '%s' has TensorFlow dtype %d" % (key, dtype))
        out = np.empty(shape, dtype=_DTYPES[dtype])
        if out.nbytes != nbytes:
            raise io_lib.IOError_(io_lib.ERR_CORRUPT, "tensor '%s': %d bytes stored for shape %s" % (key, nbytes, shape))
        io_lib.check(self._lib.fdio_bundle_read(self._handle, self._index[key], out.ctypes.data_as(ctypes.c_void_p), nbytes))
        return out

    def scalar_string(self, key: str) -> bytes:
        """A rank-0 DT_STRING tensor: varint64 length | uint32 masked crc32c of the length | bytes."""
        raw = self.raw(key)
        n, pos = _varint(raw, 0)
        return raw[pos + 4:pos + 4 + n]

    def object_graph(self) -> Optional[ObjectGraph]:
        if OBJECT_GRAPH_KEY not in self:
            return None
        return ObjectGraph.parse(self.scalar_string(OBJECT_GRAPH_KEY))


def is_tf_checkpoint(path: str) -> bool:
    return os.path.exists(path + ".index")


def _subsequence(needle: List[str], hay: List[str]) -> bool:
    it = iter(hay)
    return all(tok in it for tok in needle)


def load_variables(prefix: str, wanted: Dict[str, Tuple[int, ...]]) -> "OrderedDict[str, np.ndarray]":
    """Reads the variables named by attribute path in ``wanted`` (name -> shape).  Raises ``KeyError`` listing what could not be found
    and ``ValueError`` on a shape mismatch -- like ``load_weights`` on an incompatible checkpoint."""
    with Bundle(prefix) as bundle:
        graph = bundle.object_graph()
        value_keys = [k for k in bundle.keys if k.endswith(VARIABLE_SUFFIX) and ".OPTIMIZER_SLOT" not in k]
        tokens = {k: k[:-len(VARIABLE_SUFFIX)].split("/") for k in value_keys}
        out, missing = OrderedDict(), []
        for name, shape in wanted.items():
            path = name.split("/")
            key = graph.variable_key(path) if graph is not None else None
            if key is None or key not in bundle:
                key = name + VARIABLE_SUFFIX
            if key not in bundle:
                # same edges in the same order, possibly with wrapper edges in between; the shape must agree and the match be unique
                cands = [k for k, tk in tokens.items() if tk[-1] == path[-1] and _subsequence(path, tk) and bundle.info(k)[1] == tuple(shape)]
                key = cands[0] if len(cands) == 1 else None
            if key is None:
                missing.append(name)
                continue
            arr = bundle.tensor(key)
            if tuple(arr.shape) != tuple(shape):
                raise ValueError("Checkpoint variable %s has shape %s, the model expects %s" % (key, arr.shape, tuple(shape)))
            out[name] = arr
        if missing:
            raise KeyError("%d variables are not in checkpoint %s: %s" % (len(missing), prefix, ", ".join(missing[:8]) + (" ..." if len(missing) > 8 else "")))
        return out


def list_variables(prefix: str) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """``tf.train.list_variables``: key -> (shape, numpy dtype name)."""
    with Bundle(prefix) as bundle:
        out = OrderedDict()
        for k in bundle.keys:
            dtype, shape, _ = bundle.info(k)
            out[k] = (shape, "string" if dtype == DT_STRING else np.dtype(_DTYPES.get(dtype, np.void)).name)
        return out


def save_variables(prefix: str, weights: Dict[str, np.ndarray]):
    """Writes ``<prefix>.index``, ``<prefix>.data-00000-of-00001`` and the ``checkpoint`` state file next to them."""
    lib = io_lib.load_library()
    directory = os.path.dirname(os.path.abspath(prefix))
    os.makedirs(directory, exist_ok=True)
    # like TensorFlow, write under a temporary prefix and rename into place: a failed or interrupted save leaves the previous checkpoint
    # (best.ckpt) intact
    final_prefix, prefix = prefix, "%s.tempstate%d" % (prefix, os.getpid())
    temp_files = [prefix + ".index", prefix + ".data-00000-of-00001"]
    writer = io_lib.check_handle(lib.fdio_bundle_writer_create(prefix.encode()))
    try:
        for name, value in weights.items():
            arr = np.asarray(value, order="C")
            if arr.dtype not in _DT_OF:
                raise TypeError("variable %s has unsupported dtype %s" % (name, arr.dtype))
            dims = (ctypes.c_int64 * max(arr.ndim, 1))(*arr.shape)
            io_lib.check(lib.fdio_bundle_writer_add(writer, (name + VARIABLE_SUFFIX).encode("utf-8"), _DT_OF[arr.dtype], arr.ndim, dims,
                                                    arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
        graph = ObjectGraph.from_variable_paths(list(weights.keys())).serialize()
        io_lib.check(lib.fdio_bundle_writer_add(writer, OBJECT_GRAPH_KEY.encode(), DT_STRING, 0, None, graph, len(graph)))
    except Exception:
        lib.fdio_bundle_writer_finish(writer)  # frees the writer (there is no abort entry point); its partial files are temporaries
        for f in temp_files:
            if os.path.exists(f):
                os.remove(f)
        raise
    io_lib.check(lib.fdio_bundle_writer_finish(writer))
    os.replace(temp_files[1], final_prefix + ".data-00000-of-00001")  # data first: an index never points at a missing shard
    os.replace(temp_files[0], final_prefix + ".index")
    prefix = final_prefix
    base = os.path.basename(prefix)
    with open(os.path.join(directory, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
