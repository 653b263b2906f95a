"""ctypes binding of ``libflexdm_io.so`` (``include/flexdm_io.h``): TFRecord framing, the SequenceExample batch parser
and TensorFlow tensor-bundle checkpoints.  Host-only C++; see ``dataspec.py`` and ``checkpoint.py`` for the mirrors of the
reference interfaces built on it."""
import ctypes
import os
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflexdm_io.so")

OK = 0
ERR_ARG, ERR_IO, ERR_CORRUPT, ERR_INVALID, ERR_OOV, ERR_UNSUPPORTED, ERR_NOT_FOUND = -1, -2, -3, -4, -5, -6, -7
INT64, FLOAT32, STRING = 0, 1, 2
NONE, LOOKUP, DISCRETIZE = 0, 1, 2
OUT_INT32, OUT_FLOAT32, OUT_SPAN, OUT_SKIP = 0, 1, 2, 3

c_u8p = ctypes.POINTER(ctypes.c_uint8)


class Column(ctypes.Structure):
    _fields_ = [
        ("name", ctypes.c_char_p),
        ("is_sequence", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("transform", ctypes.c_int32),
        ("output", ctypes.c_int32),
        ("vocab_size", ctypes.c_int32),
        ("vocab_str", ctypes.POINTER(ctypes.c_char_p)),
        ("vocab_int", ctypes.POINTER(ctypes.c_int64)),
        ("num_oov_indices", ctypes.c_int32),
        ("has_mask", ctypes.c_int32),
        ("mask_str", ctypes.c_char_p),
        ("mask_int", ctypes.c_int64),
        ("n_boundaries", ctypes.c_int32),
        ("boundaries", ctypes.POINTER(ctypes.c_float)),
    ]


class PackColumn(ctypes.Structure):  # fdio_pack_column (include/flexdm_io.h)
    _fields_ = [
        ("column", ctypes.c_int32),
        ("cond_column", ctypes.c_int32),
        ("cond_mask", c_u8p),
        ("cond_n", ctypes.c_int32),
        ("rowmap", ctypes.POINTER(ctypes.c_int32)),
        ("capacity_rows", ctypes.c_int64),
        ("n_rows", ctypes.c_int64),
    ]


_SIGNATURES = {
    "fdio_last_error": (ctypes.c_char_p, []),
    "fdio_version": (ctypes.c_int, []),
    "fdio_crc32c": (ctypes.c_uint32, [ctypes.c_void_p, ctypes.c_size_t]),
    "fdio_crc32c_extend": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]),
    "fdio_crc32c_mask": (ctypes.c_uint32, [ctypes.c_uint32]),
    "fdio_crc32c_unmask": (ctypes.c_uint32, [ctypes.c_uint32]),
    "fdio_tfrecord_open": (ctypes.c_void_p, [ctypes.c_char_p, ctypes.c_int]),
    "fdio_tfrecord_close": (None, [ctypes.c_void_p]),
    "fdio_tfrecord_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "fdio_tfrecord_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]),
    "fdio_tfrecord_write": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int64]),
    "fdio_schema_create": (ctypes.c_void_p, [ctypes.POINTER(Column), ctypes.c_int32]),
    "fdio_schema_destroy": (None, [ctypes.c_void_p]),
    "fdio_batch_steps": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                        ctypes.POINTER(ctypes.c_int32), ctypes.c_int32]),
    "fdio_parse_batch": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                        ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32]),
    "fdio_parse_batch_packed": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32,
                                               ctypes.c_int32, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int32, ctypes.POINTER(PackColumn), ctypes.c_int32,
                                               ctypes.c_int32]),
    "fdio_bundle_open": (ctypes.c_void_p, [ctypes.c_char_p]),
    "fdio_bundle_close": (None, [ctypes.c_void_p]),
    "fdio_bundle_count": (ctypes.c_int32, [ctypes.c_void_p]),
    "fdio_bundle_key": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int32]),
    "fdio_bundle_find": (ctypes.c_int32, [ctypes.c_void_p, ctypes.c_char_p]),
    "fdio_bundle_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]),
    "fdio_bundle_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]),
    "fdio_bundle_writer_create": (ctypes.c_void_p, [ctypes.c_char_p]),
    "fdio_bundle_writer_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64),
                                              ctypes.c_void_p, ctypes.c_int64]),
    "fdio_bundle_writer_finish": (ctypes.c_int, [ctypes.c_void_p]),
}

_lib = None


class IOError_(RuntimeError):
    """An error reported by libflexdm_io (code + message of ``fdio_last_error``)."""

    def __init__(self, code: int, message: str):
        super().__init__("%s (fdio code %d)" % (message, code))
        self.code = code


class InvalidArgumentError(IOError_, ValueError):
    """Stands where ``tf.errors.InvalidArgumentError`` does in the reference (parse and lookup failures)."""


def exported_symbols() -> List[str]:
    return sorted(_SIGNATURES)


def load_library():
    """Loads the library next to this file; there is no Python fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (make -C flex_dm_b200/csrc)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load_library().fdio_last_error().decode("utf-8", "replace")


def check(code: int):
    if code != OK:
        cls = InvalidArgumentError if code in (ERR_INVALID, ERR_OOV, ERR_CORRUPT) else IOError_
        raise cls(code, last_error())


def check_handle(handle):
    if not handle:
        msg = last_error()
        raise (FileNotFoundError(msg) if "cannot open" in msg else IOError_(ERR_IO, msg))
    return handle


def crc32c(data: bytes) -> int:
    return load_library().fdio_crc32c(data, len(data))


def masked_crc32c(data: bytes) -> int:
    lib = load_library()
    return lib.fdio_crc32c_mask(lib.fdio_crc32c(data, len(data)))
