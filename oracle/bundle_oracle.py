"""TEST INFRASTRUCTURE ONLY -- pure-Python restatement of TensorFlow's tensor-bundle checkpoint format, the checker of
``flex_dm_b200/csrc/io/bundle.cc`` (reference call sites: ``train.py:67-69,94-97``, ``eval.py:169-172``, ``helpers/callbacks.py:49-56``).

Formats restated from their published descriptions (TensorFlow is not under /root/reference; **parity unpinned**: no TF-written file is
available offline): the LevelDB table format used by ``tensorflow/core/lib/io/table`` (``doc/table_format.md``) and
``tensorflow/core/protobuf/tensor_bundle.proto``.  Written independently of the C++ (struct + bitwise CRC from ``dataspec_oracle``)."""
import struct
from typing import Dict, List, Tuple

import numpy as np

from .dataspec_oracle import _fields, _varint, crc32c, mask_crc

MAGIC = 0xDB4775248B80FB57
NP_OF_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}
DT_OF_NP = {np.dtype(v): k for k, v in NP_OF_DT.items()}


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _block(data: bytes, offset: int, size: int) -> List[Tuple[bytes, bytes]]:
    body = data[offset:offset + size]
    ctype = data[offset + size]
    (stored,) = struct.unpack_from("<I", data, offset + size + 1)
    assert mask_crc(crc32c(body + bytes([ctype]))) == stored, "block crc"
    assert ctype == 0, "compressed block"
    (n_restarts,) = struct.unpack_from("<I", body, len(body) - 4)
    limit = len(body) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < limit:
        shared, pos = _varint(body, pos)
        non_shared, pos = _varint(body, pos)
        vlen, pos = _varint(body, pos)
        key = key[:shared] + body[pos:pos + non_shared]
        pos += non_shared
        out.append((key, body[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path: str) -> List[Tuple[bytes, bytes]]:
    data = open(path, "rb").read()
    footer = data[-48:]
    assert struct.unpack("<Q", footer[40:])[0] == MAGIC, "magic"
    pos = 0
    _, pos = _varint(footer, pos)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = []
    for _, handle in _block(data, idx_off, idx_size):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        out += _block(data, off, size)
    return out


def read_bundle(prefix: str) -> Dict[str, np.ndarray]:
    """Every non-string tensor of the bundle (single shard)."""
    entries = read_table(prefix + ".index")
    assert entries[0][0] == b"", "header first"
    shard = open(prefix + ".data-00000-of-00001", "rb").read()
    out = {}
    for key, value in entries[1:]:
        dtype, dims, offset, size, crc = 0, [], 0, 0, 0
        for field, wt, v in _fields(value):
            if field == 1:
                dtype = v
            elif field == 2:
                for f2, _, dim in _fields(v):
                    if f2 == 2:
                        d = 0
                        for f3, _, v3 in _fields(dim):
                            if f3 == 1:
                                d = v3
                        dims.append(d)
            elif field == 4:
                offset = v
            elif field == 5:
                size = v
            elif field == 6:
                (crc,) = struct.unpack("<I", v)
        raw = shard[offset:offset + size]
        if dtype == 7:
            continue
        assert mask_crc(crc32c(raw)) == crc, "tensor crc of %r" % key
        out[key.decode()] = np.frombuffer(raw, dtype=NP_OF_DT[dtype]).reshape(dims)
    return out


def _emit_block(out: bytearray, entries: List[Tuple[bytes, bytes]], restart_interval: int) -> bytes:
    body, restarts, last = bytearray(), [], b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(last), len(key)) and last[shared] == key[shared]:
                shared += 1
        body += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    offset = len(out)
    out += body + b"\0" + struct.pack("<I", mask_crc(crc32c(bytes(body) + b"\0")))
    return _put_varint(offset) + _put_varint(len(body))


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], entries_per_block: int = 5, restart_interval: int = 3):
    """An independent writer (different block and restart sizes from the C++ one) producing files the C++ reader must accept."""
    shard = bytearray()
    table: List[Tuple[bytes, bytes]] = [(b"", b"\x08\x01\x1a\x02\x08\x01")]  # num_shards = 1, version { producer = 1 }
    for key in sorted(tensors, key=lambda k: k.encode()):
        arr = np.asarray(tensors[key], order="C")
        raw = arr.tobytes()
        shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in [(b"\x08" + _put_varint(s)) if s else b"" for s in arr.shape])
        value = b"\x08" + _put_varint(DT_OF_NP[arr.dtype]) + b"\x12" + _put_varint(len(shape)) + shape
        if len(shard):
            value += b"\x20" + _put_varint(len(shard))
        if len(raw):
            value += b"\x28" + _put_varint(len(raw))
        value += b"\x35" + struct.pack("<I", mask_crc(crc32c(raw)))
        shard += raw
        table.append((key.encode(), value))
    out = bytearray()
    index = []
    for i in range(0, len(table), entries_per_block):
        chunk = table[i:i + entries_per_block]
        index.append((chunk[-1][0], _emit_block(out, chunk, restart_interval)))
    meta = _emit_block(out, [], restart_interval)
    idx = _emit_block(out, index, restart_interval)
    footer = (meta + idx).ljust(40, b"\0") + struct.pack("<Q", MAGIC)
    open(prefix + ".index", "wb").write(bytes(out) + footer)
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(shard))
