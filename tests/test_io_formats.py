"""CPU tests of the data formats either side of the MFP step (``include/flexdm_io.h``):

* TFRecord framing and CRC-32C against RFC 3720 test vectors and the pure-Python oracle;
* the native SequenceExample batch parser (``DataSpec.parse_fn``, reference ``data/spec.py:255-287``) against ``oracle/dataspec_oracle.py``
  on records serialised by **google.protobuf** (an encoder independent of both), bit-exact, including the edge cases the format has:
  ragged and empty documents, unpacked repeated fields, out-of-vocabulary tokens, missing / mistyped / mis-sized features;
* ``DataSpec`` as the reference's callers use it (``train.py:38-52``): input columns, datasets, shuffle / repeat / batch, ``unbatch``;
* TensorFlow tensor-bundle checkpoints against ``oracle/bundle_oracle.py`` in both directions, object-graph restore, corruption.
"""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

from flex_dm_b200 import checkpoint, io_lib
from flex_dm_b200.dataspec import BUILTIN_SPECS, DataSpec, IntegerLookup, SequenceDiscretizer, StringLookup, TFRecordFile, encode_feature, encode_sequence_example, write_tfrecord
from flex_dm_b200.spec import make_input_columns
from flex_dm_b200.synthetic import synthetic_vocabulary, write_synthetic_dataset
from oracle import bundle_oracle, dataspec_oracle as DO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------------------------------ C ABI
def _declared():
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "flexdm_io.h")).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(fdio_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_and_binding_cover_the_header():
    lib = io_lib.load_library()
    names = _declared()
    assert "fdio_parse_batch" in names and "fdio_bundle_open" in names and "fdio_tfrecord_open" in names
    for n in names:
        assert hasattr(lib, n), n
    assert io_lib.exported_symbols() == names
    assert lib.fdio_version() == 1


# ------------------------------------------------------------------------------------------------------------------------ CRC-32C
def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 and the classic check value
    assert io_lib.crc32c(b"123456789") == 0xE3069283
    assert io_lib.crc32c(bytes(32)) == 0x8A9136AA
    assert io_lib.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert io_lib.crc32c(bytes(range(32))) == 0x46DD794E
    assert io_lib.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert io_lib.crc32c(b"") == 0
    rng = np.random.default_rng(0)
    lib = io_lib.load_library()
    for n in (1, 7, 8, 9, 63, 1000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert io_lib.crc32c(data) == DO.crc32c(data)
        assert io_lib.masked_crc32c(data) == DO.mask_crc(DO.crc32c(data))
        cut = n // 3
        assert lib.fdio_crc32c_extend(io_lib.crc32c(data[:cut]), data[cut:], n - cut) == io_lib.crc32c(data)
        assert lib.fdio_crc32c_unmask(io_lib.masked_crc32c(data)) == io_lib.crc32c(data)


# ------------------------------------------------------------------------------------------------------------------------ TFRecord
def test_tfrecord_framing_both_directions(tmp_path):
    rng = np.random.default_rng(1)
    records = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (0, 1, 5, 4096, 70000)]
    ours = str(tmp_path / "ours.tfrecord")
    write_tfrecord(ours, records)
    assert open(ours, "rb").read() == b"".join(DO.frame_record(r) for r in records)
    assert DO.read_tfrecord(ours) == records
    theirs = str(tmp_path / "theirs.tfrecord")
    open(theirs, "wb").write(b"".join(DO.frame_record(r) for r in records))
    f = TFRecordFile(theirs, verify_crc=2)
    assert len(f) == len(records) and [f.record(i) for i in range(len(f))] == records
    empty = str(tmp_path / "empty.tfrecord")
    open(empty, "wb").close()
    assert len(TFRecordFile(empty)) == 0


def test_tfrecord_corruption_is_detected(tmp_path):
    blob = bytearray(DO.frame_record(b"hello world") + DO.frame_record(b"second"))
    good = str(tmp_path / "good.tfrecord")
    open(good, "wb").write(blob)
    assert len(TFRecordFile(good)) == 2
    flipped = bytearray(blob)
    flipped[14] ^= 1  # payload byte
    p = str(tmp_path / "payload.tfrecord")
    open(p, "wb").write(flipped)
    with pytest.raises(io_lib.IOError_, match="corrupted record data"):
        TFRecordFile(p, verify_crc=2)
    assert len(TFRecordFile(p, verify_crc=1)) == 2  # length CRCs only
    flipped = bytearray(blob)
    flipped[0] ^= 1  # length
    open(p, "wb").write(flipped)
    with pytest.raises(io_lib.IOError_, match="corrupted record length"):
        TFRecordFile(p, verify_crc=1)
    open(p, "wb").write(blob[:-3])
    with pytest.raises(io_lib.IOError_, match="truncated"):
        TFRecordFile(p)
    with pytest.raises(FileNotFoundError):
        TFRecordFile(str(tmp_path / "missing.tfrecord"))


# ------------------------------------------------------------------------------------------------------------------------ protobuf as the independent encoder
@pytest.fixture(scope="module")
def pb():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto(name="example_for_tests.proto", package="tensorflow", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def message(name, *fields):
        m = fd.message_type.add(name=name)
        for fname, number, ftype, label, type_name, oneof in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = ".tensorflow." + type_name
            if oneof is not None:
                f.oneof_index = oneof
        return m

    message("BytesList", ("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None, None))
    message("FloatList", ("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None, None))
    message("Int64List", ("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None, None))
    feature = message("Feature", ("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "BytesList", 0),
                      ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "FloatList", 0),
                      ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Int64List", 0))
    feature.oneof_decl.add(name="kind")
    features = message("Features", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, "Features.FeatureEntry", None))
    entry = features.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow.Feature")
    message("FeatureList", ("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, "Feature", None))
    lists = message("FeatureLists", ("feature_list", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, "FeatureLists.FeatureListEntry", None))
    entry = lists.nested_type.add(name="FeatureListEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".tensorflow.FeatureList")
    message("SequenceExample", ("context", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "Features", None),
            ("feature_lists", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "FeatureLists", None))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.SequenceExample"))


def _fill(feature, dtype, values):
    if dtype == "string":
        feature.bytes_list.value.extend(v if isinstance(v, bytes) else v.encode() for v in values)
    elif dtype == "float32":
        feature.float_list.value.extend(float(v) for v in values)
    else:
        feature.int64_list.value.extend(int(v) for v in values)


def _random_documents(pb, columns, vocabulary, n_docs, rng, max_len=9, oov=True):
    """Raw-valued random documents serialised by google.protobuf; includes one empty and one single-element document."""
    pre = DO.make_preprocessors(columns, vocabulary)
    records = []
    for d in range(n_docs):
        n = 0 if d == 1 else 1 if d == 2 else int(rng.integers(1, max_len + 1))
        ex = pb()
        for name, column in columns.items():
            width = int(np.prod(column.get("shape", (1,))))
            dtype = column["dtype"]
            lookup = pre.get(name) if "lookup" in column else None

            def draw():
                if name == "length":
                    return [max(n, 1)]
                if lookup is not None:
                    first = (0 if lookup.mask is None else 1) + lookup.num_oov
                    pool = lookup.tokens[first:] + (["never seen"] if (oov and lookup.num_oov and dtype == "string") else [])
                    if lookup.mask is not None:
                        pool = pool + [lookup.mask]
                    return [pool[int(rng.integers(len(pool)))] for _ in range(width)]
                if dtype == "string":
                    return ["s%d" % rng.integers(1000) for _ in range(width)]
                if dtype == "float32":
                    # include exact boundaries of the discretiser and the ends of the range
                    special = [0.0, 1.0, 1 / 63, 62 / 63, 0.5, 1 / 7]
                    return [special[int(rng.integers(len(special)))] if rng.random() < 0.3 else float(np.float32(rng.random())) for _ in range(width)]
                return [int(rng.integers(0, 256 if "discretize" in column else column.get("max", 1) + 1)) for _ in range(width)]

            if column.get("is_sequence"):
                fl = ex.feature_lists.feature_list[name]
                for _ in range(n):
                    _fill(fl.feature.add(), dtype, draw())
            else:
                _fill(ex.context.feature[name], dtype, draw())
        records.append(ex.SerializeToString())
    return records


def _assert_batches_equal(got, want):
    assert list(got.keys()) == list(want.keys())
    for k in want:
        g = got[k].numpy() if hasattr(got[k], "numpy") else got[k]
        assert g.shape == want[k].shape, (k, g.shape, want[k].shape)
        if want[k].dtype == object:
            assert g.tolist() == want[k].tolist(), k
        else:
            assert g.dtype == want[k].dtype, (k, g.dtype)
            assert np.array_equal(g, want[k]), k


@pytest.mark.parametrize("name", ["crello", "rico"])
@pytest.mark.parametrize("threads", [1, 3])
def test_parse_fn_matches_the_oracle_on_protobuf_serialised_records(pb, tmp_path, name, threads):
    columns = BUILTIN_SPECS[name]["columns"]
    vocabulary = synthetic_vocabulary(name, make_input_columns(name))
    (tmp_path / "vocabulary.json").write_text(__import__("json").dumps(vocabulary))
    (tmp_path / "count.json").write_text('{"train": 0}')
    spec = DataSpec(name, str(tmp_path), num_threads=threads)
    rng = np.random.default_rng(5)
    records = _random_documents(pb, columns, vocabulary, 11, rng)
    got = spec.parse_fn(records)
    want = DO.parse_fn(columns, vocabulary, records)
    _assert_batches_equal(got, want)
    assert got["left"].shape[1] == max(len(DO.decode_sequence_example(r)[1]["left"]) for r in records)  # padded to the batch maximum
    # fixed-shape batches: same values, more padding
    padded = spec.parse_fn(records, pad_to=16)
    for k, v in want.items():
        if columns[k].get("is_sequence") and v.dtype != object:
            assert np.array_equal(padded[k].numpy()[:, : v.shape[1]], v), k
            assert (padded[k].numpy()[:, v.shape[1]:] == padded[k].numpy()[1, 0]).all(), k  # document 1 is empty: all padding
    with pytest.raises(ValueError, match="pad_to"):
        spec.parse_fn(records, pad_to=2)
    # an empty batch and a batch of empty documents
    assert spec.parse_fn([])["left"].shape == (0, 0, 1)
    assert spec.parse_fn([records[1]])["left"].shape == (1, 0, 1)


def test_in_house_encoder_is_read_back_by_protobuf(pb):
    ctx = {"length": encode_feature([3], "int64"), "id": encode_feature([b"abc"], "string")}
    fl = {"left": [encode_feature([0.25], "float32"), encode_feature([0.5], "float32")], "color": [encode_feature([1, 255, -7], "int64")] * 2,
          "empty": []}
    ex = pb.FromString(encode_sequence_example(ctx, fl))
    assert list(ex.context.feature["length"].int64_list.value) == [3]
    assert list(ex.context.feature["id"].bytes_list.value) == [b"abc"]
    assert [list(f.float_list.value) for f in ex.feature_lists.feature_list["left"].feature] == [[0.25], [0.5]]
    assert [list(f.int64_list.value) for f in ex.feature_lists.feature_list["color"].feature] == [[1, 255, -7]] * 2
    assert len(ex.feature_lists.feature_list["empty"].feature) == 0


def _tiny_spec(tmp_path, columns, vocabulary=None):
    import yaml

    (tmp_path / "vocabulary.json").write_text(__import__("json").dumps(vocabulary or {}))
    (tmp_path / "count.json").write_text('{"train": 0}')
    path = tmp_path / "tiny-spec.yml"
    path.write_text(yaml.safe_dump({"name": "tiny", "columns": columns}, sort_keys=False))
    return DataSpec(str(path), str(tmp_path), num_threads=2)


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def test_unpacked_repeated_fields_and_negative_integers(tmp_path):
    """Parsers must accept both encodings of repeated scalars; int64 values are two's-complement varints."""
    spec = _tiny_spec(tmp_path, {"n": {"dtype": "int64"}, "x": {"dtype": "float32", "is_sequence": True, "shape": [2]},
                                 "k": {"dtype": "int64", "is_sequence": True, "shape": [2]}})
    unpacked_floats = _ld(2, b"".join(b"\x0d" + struct.pack("<f", v) for v in (1.5, -2.0)))  # field 1, wire type 5, twice
    unpacked_ints = _ld(3, b"\x08" + _varint(7) + b"\x08" + _varint((1 << 64) - 3))  # 7, -3
    rec = encode_sequence_example({"n": _ld(3, b"\x08" + _varint((1 << 64) - 1))}, {"x": [unpacked_floats], "k": [unpacked_ints]})
    out = spec.parse_fn([rec])
    assert out["n"].tolist() == [[-1]]
    assert out["x"].tolist() == [[[1.5, -2.0]]]
    assert out["k"].tolist() == [[[7, -3]]]
    want = DO.parse_fn(spec.columns, {}, [rec])
    _assert_batches_equal(out, want)


def test_parse_errors_name_the_record_and_the_key(tmp_path):
    cols = {"n": {"dtype": "int64"}, "t": {"dtype": "string", "is_sequence": True, "lookup": {"num_oov_indices": 0, "mask_token": ""}},
            "x": {"dtype": "float32", "is_sequence": True, "shape": [2]}}
    spec = _tiny_spec(tmp_path, cols, {"t": ["a", "b"]})
    f = encode_feature
    ok = encode_sequence_example({"n": f([1], "int64")}, {"t": [f(["a"], "string")], "x": [f([0.0, 1.0], "float32")]})
    assert spec.parse_fn([ok])["t"].tolist() == [[[1]]]  # index 0 is the mask token ""
    cases = [
        (encode_sequence_example({}, {"t": [], "x": []}), "key 'n': feature is required"),
        (encode_sequence_example({"n": f([1], "int64")}, {"t": []}), "key 'x': feature list is required"),
        (encode_sequence_example({"n": f([1.0], "float32")}, {"t": [], "x": []}), "key 'n': feature kind does not match"),
        (encode_sequence_example({"n": f([1, 2], "int64")}, {"t": [], "x": []}), "values size 2 but output shape holds 1"),
        (encode_sequence_example({"n": f([1], "int64")}, {"t": [], "x": [f([0.0, 1.0], "float32"), f([0.0], "float32")]}), "key 'x', index 1: number of float values"),
        (encode_sequence_example({"n": f([1], "int64")}, {"t": [f(["zzz"], "string")], "x": [f([0.0, 1.0], "float32")]}), "key 't', index 0: value is not in the lookup vocabulary"),
        (encode_sequence_example({"n": f([1], "int64")}, {"t": [b""], "x": [f([0.0, 1.0], "float32")]}), "key 't', index 0: feature holds no values"),
        (ok[:-3], "malformed"),
    ]
    for rec, message in cases:
        with pytest.raises(io_lib.InvalidArgumentError, match=re.escape(message)):
            spec.parse_fn([ok, rec])
        with pytest.raises(ValueError, match="record 1"):  # InvalidArgumentError is a ValueError; the failing record is named
            spec.parse_fn([ok, rec])
    # padding a string lookup that has neither a mask token nor an OOV index for "" is an error only when padding happens
    cols2 = {"t": {"dtype": "string", "is_sequence": True, "lookup": {"num_oov_indices": 0, "mask_token": None}}}
    (tmp_path / "b").mkdir()
    spec2 = _tiny_spec(tmp_path / "b", cols2, {"t": ["a"]})
    one = encode_sequence_example({}, {"t": [f(["a"], "string")]})
    two = encode_sequence_example({}, {"t": [f(["a"], "string")] * 2})
    assert spec2.parse_fn([one, one])["t"].tolist() == [[[0]], [[0]]]
    with pytest.raises(io_lib.InvalidArgumentError, match="padding value"):
        spec2.parse_fn([one, two])


def test_mutated_records_never_crash_the_parser(pb, tmp_path):
    """Every length and offset in a record is attacker-controlled: random byte flips, truncations and splices must end in a clean
    InvalidArgumentError or a successful parse, never in an out-of-bounds access."""
    columns = BUILTIN_SPECS["rico"]["columns"]
    vocabulary = synthetic_vocabulary("rico", make_input_columns("rico"))
    (tmp_path / "vocabulary.json").write_text(__import__("json").dumps(vocabulary))
    (tmp_path / "count.json").write_text('{"train": 0}')
    spec = DataSpec("rico", str(tmp_path), num_threads=2)
    rng = np.random.default_rng(11)
    records = _random_documents(pb, columns, vocabulary, 6, rng)
    outcomes = {"ok": 0, "error": 0}
    for trial in range(1500):
        rec = bytearray(records[trial % len(records)])
        if not rec:
            continue
        kind = trial % 3
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                rec[int(rng.integers(len(rec)))] = int(rng.integers(256))
        elif kind == 1:
            rec = rec[: int(rng.integers(len(rec)))]
        else:
            a, b = sorted(int(x) for x in rng.integers(0, len(rec), 2))
            rec = rec[:a] + rec[b:]
        try:
            spec.parse_fn([records[0], bytes(rec)], pad_to=64)
            outcomes["ok"] += 1
        except (io_lib.InvalidArgumentError, ValueError):
            outcomes["error"] += 1
    assert outcomes["error"] > 100 and outcomes["ok"] + outcomes["error"] >= 1400


def test_lookup_and_discretizer_semantics():
    s = StringLookup(vocabulary=["x", "y"], num_oov_indices=1, mask_token="")
    assert s.get_vocabulary() == ["", "[UNK]", "x", "y"] and s.vocabulary_size() == 4
    assert [int(s(v)) for v in ("", "x", "y", "q", b"y")] == [0, 2, 3, 1, 3]
    s0 = StringLookup(vocabulary=["x", "y"], num_oov_indices=0, mask_token=None)
    assert s0.get_vocabulary() == ["x", "y"] and int(s0("y")) == 1
    with pytest.raises(ValueError):
        s0("q")
    i = IntegerLookup(vocabulary=list(range(1, 51)), num_oov_indices=0, mask_value=None)  # the reference's `length` column
    assert i.vocabulary_size() == 50 and int(i(1)) == 0 and int(i(50)) == 49
    with pytest.raises(ValueError):
        StringLookup(vocabulary=["a", "a"])
    with pytest.raises(NotImplementedError):
        StringLookup(vocabulary=["a"], num_oov_indices=2)
    # Bucketize: number of float32 boundaries <= float32(x); boundaries = linspace(min, max, bins)[1:] (spec.py:97-103)
    d = SequenceDiscretizer(list(np.linspace(0.0, 1.0, 64))[1:])
    assert len(d.bin_boundaries) + 1 == 64
    assert d([0.0, 1 / 63, np.float32(1 / 63), 1.0, 2.0, -1.0, 0.999]).tolist() == [0, 1, 1, 63, 63, 0, 62]
    c = SequenceDiscretizer(list(np.linspace(0, 255, 16))[1:])
    assert c([0, 16, 17, 254, 255]).tolist() == [0, 0, 1, 14, 15]


# ------------------------------------------------------------------------------------------------------------------------ DataSpec as train.py uses it
@pytest.fixture(scope="module")
def crello_dir(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("crello"))
    written = write_synthetic_dataset(root, "crello", {"train": 37, "val": 9, "test": 0}, seq_len=14, shards=3, seed=3)
    return root, written


def test_dataspec_input_columns_sizes_and_steps(crello_dir):
    root, _ = crello_dir
    spec = DataSpec("crello", root, batch_size=8)
    cols = spec.make_input_columns()
    want = make_input_columns("crello")
    assert list(cols.keys()) == list(want.keys())
    for k in want:
        assert dict(cols[k]) == dict(want[k]), k
    assert spec.size("train") == 37 and spec.steps_per_epoch("train") == 5 and spec.steps_per_epoch("val", 4) == 3
    assert cols["type"]["primary_label"] == 0 and cols["font_family"]["input_dim"] == 35  # the rare font is filtered by min_freq
    with pytest.raises(AssertionError):
        spec.make_dataset("nope")
    with pytest.raises(FileNotFoundError):
        DataSpec("no-such-dataset", root)


def test_dataset_batches_reproduce_the_exported_documents(crello_dir):
    root, written = crello_dir
    spec = DataSpec("crello", root, batch_size=8)
    batches = list(spec.make_dataset("train", shuffle=False, strings=True, verify_crc=2))
    assert [b["length"].shape[0] for b in batches] == [8, 8, 8, 8, 5]  # the last batch is partial (no drop_remainder)
    docs = {k: np.concatenate([w[k] for w in written["train"]]) for k in written["train"][0]}
    row = 0
    for b in batches:
        n = b["length"].shape[0]
        S = b["left"].shape[1]
        assert S == int(b["length"].max()) + 1  # padded to the batch's longest document
        for k, v in docs.items():
            want = v[row:row + n]
            if want.ndim == 3:
                want = want[:, :S]
            assert np.array_equal(b[k].numpy(), want), k
            assert b[k].dtype == (__import__("torch").float32 if want.dtype == np.float32 else __import__("torch").int32)
        assert b["id"].shape == (n, 1) and b["uuid"].shape == (n, S, 1) and b["id"][0, 0].startswith(b"train-")
        row += n
    assert row == 37
    # demo-only byte strings are validated but not emitted by default
    assert "id" not in next(iter(spec.make_dataset("train")))


def test_packed_batches_equal_pack_batch_of_the_dense_ones(crello_dir, tmp_path):
    """``make_dataset(packed=True)`` writes the numerical columns packed straight from the records (``fdio_parse_batch_packed``): every
    batch must be bit-identical to ``pack_batch`` of the dense batch -- the rows of the elements that carry the field (valid position and
    type gate, data/crello-spec.yml:88-121), document by document, and the element -> row map -- for ragged batches, fixed-shape ones
    (``pad_to``), one parser thread and many; every other column is unchanged."""
    from flex_dm_b200.data import ROWS_SUFFIX, pack_batch

    root, _ = crello_dir
    for threads, pad_to in ((1, None), (5, None), (3, 20)):
        spec = DataSpec("crello", root, batch_size=8, num_threads=threads)
        cols = spec.make_input_columns()
        dense = list(spec.make_dataset("train", shuffle=False, pad_to=pad_to))
        packed = list(spec.make_dataset("train", shuffle=False, pad_to=pad_to, packed=True))
        assert len(dense) == len(packed) == 5
        n_packed_keys = 0
        for d, p in zip(dense, packed):
            want = pack_batch({k: v.numpy() for k, v in d.items()}, cols)
            assert set(want.keys()) == set(p.keys())
            for k, v in want.items():
                got = p[k].numpy()
                assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
                n_packed_keys += k.endswith(ROWS_SUFFIX)
            for k in ("image_embedding", "text_embedding"):
                assert p[k].shape[0] < d[k].shape[0] * d[k].shape[1]  # fewer rows than elements: the type gate and the padding
        assert n_packed_keys == 2 * 5
    # a document with more elements than pad_to is reported like by the dense parse
    spec = DataSpec("crello", root, batch_size=8)
    with pytest.raises(ValueError):
        next(iter(spec.make_dataset("train", shuffle=False, pad_to=3, packed=True)))


def test_parser_worker_pool_survives_a_fork(crello_dir):
    """The parser keeps a pool of worker threads between calls; a forked child (multiprocessing's default start method) has none of the
    parent's threads and must start its own instead of waiting for workers that do not exist."""
    import os

    root, _ = crello_dir
    spec = DataSpec("crello", root, batch_size=8, num_threads=4)
    want = next(iter(spec.make_dataset("train", shuffle=False, packed=True)))
    pid = os.fork()
    if pid == 0:
        code = 3
        try:
            got = next(iter(spec.make_dataset("train", shuffle=False, packed=True)))
            code = 0 if all(np.array_equal(np.asarray(got[k]), np.asarray(want[k])) for k in want) else 4
        finally:
            os._exit(code)
    deadline = __import__("time").time() + 60
    while True:
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        if __import__("time").time() > deadline:
            os.kill(pid, 9)
            pytest.fail("the forked child hung in the parser")
        __import__("time").sleep(0.05)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0


def test_dataset_shuffle_repeat_and_prefetch(crello_dir):
    root, _ = crello_dir
    spec = DataSpec("crello", root, batch_size=8)
    plain = [b["left"].numpy() for b in spec.make_dataset("train", shuffle=False, pad_to=14, prefetch=0)]
    ids = lambda ds, n=None: [bytes(x) for i, b in zip(range(n or 10 ** 9), ds) for x in b["id"][:, 0]]
    base = ids(spec.make_dataset("train", shuffle=False, strings=True))
    a = ids(spec.make_dataset("train", shuffle=True, strings=True, seed=1))
    a2 = ids(spec.make_dataset("train", shuffle=True, strings=True, seed=1))
    b = ids(spec.make_dataset("train", shuffle=True, strings=True, seed=2))
    assert sorted(a) == sorted(base) and a != base and a == a2 and a != b  # a permutation, reproducible per seed
    ds = spec.make_dataset("train", shuffle=True, strings=True, seed=1)
    assert ids(ds) != ids(ds)  # reshuffled each pass
    windowed = ids(spec.make_dataset("train", shuffle=5, strings=True, seed=4))
    assert sorted(windowed) == sorted(base) and windowed != base
    assert all(base.index(x) <= i + 5 for i, x in enumerate(windowed))  # a buffer of 5 cannot pull an item forward by more than 5
    rep = ids(spec.make_dataset("train", shuffle=False, repeat=True, strings=True), n=12)  # batches straddle the pass boundary
    assert len(rep) == 96 and rep[:37] == base and rep[37:74] == base
    pre = [b["left"].numpy() for b in spec.make_dataset("train", shuffle=False, pad_to=14, prefetch=3)]
    assert len(pre) == len(plain) and all(np.array_equal(x, y) for x, y in zip(pre, plain))


def test_prefetch_surfaces_parse_errors(tmp_path):
    cols = {"n": {"dtype": "int64"}}
    spec = _tiny_spec(tmp_path, cols)
    (tmp_path / "count.json").write_text('{"train": 2}')
    write_tfrecord(str(tmp_path / "train-0.tfrecord"), [encode_sequence_example({"n": encode_feature([1], "int64")}, {}), b"\x0a\x05abc"])
    spec = DataSpec(str(tmp_path / "tiny-spec.yml"), str(tmp_path), batch_size=1)
    it = iter(spec.make_dataset("train", prefetch=2))
    assert next(it)["n"].tolist() == [[1]]
    with pytest.raises(io_lib.InvalidArgumentError):
        next(it)


def test_dataset_iterator_answers_get_next(crello_dir):
    """eval.py:47-49 drives the dataset with ``iterator = iter(dataset)`` / ``iterator.get_next()``: same batches as ``next``; the end
    of a non-repeating split is ``StopIteration``."""
    root, written = crello_dir
    spec = DataSpec("crello", root, batch_size=4)
    for prefetch in (0, 2):
        a = iter(spec.make_dataset("val", shuffle=False, prefetch=prefetch))
        b = iter(spec.make_dataset("val", shuffle=False, prefetch=prefetch))
        steps = spec.steps_per_epoch("val")
        for _ in range(steps):
            x, y = a.get_next(), next(b)
            assert list(x) == list(y) and all(np.array_equal(np.asarray(x[k]), np.asarray(y[k])) for k in x)
        with pytest.raises(StopIteration):
            a.get_next()
        a.close()


def test_unbatch_undoes_lookup_and_discretisation(crello_dir):
    root, written = crello_dir
    spec = DataSpec("crello", root, batch_size=4)
    batch = next(iter(spec.make_dataset("val", shuffle=False, strings=True)))
    items = spec.unbatch(batch)
    assert len(items) == 4
    src = {k: np.concatenate([w[k] for w in written["val"]]) for k in written["val"][0]}  # shard-major document order
    vocab = spec.preprocessor["type"].get_vocabulary()
    for i, item in enumerate(items):
        n = int(src["length"][i, 0]) + 1
        assert len(item["elements"]) == n and item["length"] == n
        for j, e in enumerate(item["elements"]):
            assert e["type"] == vocab[src["type"][i, j, 0]]
            assert e["left"] == pytest.approx(src["left"][i, j, 0] / 63.0)
            assert e["color"] == pytest.approx([17.0 * c for c in src["color"][i, j]])
            assert len(e["image_embedding"]) == 512
    # logits (one more axis) are turned into labels first (spec.py:289-299)
    import torch

    logits = torch.nn.functional.one_hot(batch["left"].long(), 64).float()
    again = spec.unbatch(dict(batch, left=logits))
    assert again[0]["elements"][0]["left"] == items[0]["elements"][0]["left"]


def test_rico_spec_without_vocabulary_range_columns(tmp_path):
    written = write_synthetic_dataset(str(tmp_path), "rico", {"train": 6}, seq_len=7, shards=1, seed=2)
    spec = DataSpec("rico", str(tmp_path), batch_size=6)
    cols = spec.make_input_columns()
    want = make_input_columns("rico")
    for k in want:
        assert dict(cols[k]) == dict(want[k]), k
    batch = next(iter(spec.make_dataset("train")))
    for k, v in written["train"][0].items():
        assert np.array_equal(batch[k].numpy(), v[:, : batch["left"].shape[1]] if v.ndim == 3 else v), k


# ------------------------------------------------------------------------------------------------------------------------ checkpoints
def _weights(rng, n=40):
    w = {}
    for i in range(n):
        w["model/blocks/seq2seq/seq2seq_%d/attn/dense_query/kernel" % i] = rng.standard_normal((6, 4)).astype(np.float32)
        w["model/blocks/seq2seq/seq2seq_%d/attn/dense_query/bias" % i] = rng.standard_normal((4,)).astype(np.float32)
    w["model/encoder/input_layer/left/embeddings"] = rng.standard_normal((66, 8)).astype(np.float32)
    return w


def test_bundle_written_by_the_oracle_is_read_natively(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {k + checkpoint.VARIABLE_SUFFIX: v for k, v in _weights(rng).items()}
    tensors["save_counter" + checkpoint.VARIABLE_SUFFIX] = np.asarray(3, dtype=np.int64)
    tensors["empty" + checkpoint.VARIABLE_SUFFIX] = np.zeros((0, 4), dtype=np.float32)
    prefix = str(tmp_path / "best.ckpt")
    bundle_oracle.write_bundle(prefix, tensors, entries_per_block=5, restart_interval=3)
    with checkpoint.Bundle(prefix) as b:
        assert b.keys == sorted(tensors, key=lambda k: k.encode())
        for k, v in tensors.items():
            got = b.tensor(k)
            assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
        assert b.object_graph() is None
    # no object graph: variables are matched by key
    wanted = {k[: -len(checkpoint.VARIABLE_SUFFIX)]: v.shape for k, v in tensors.items() if k.startswith("model/")}
    got = checkpoint.load_variables(prefix, wanted)
    assert all(np.array_equal(got[k], tensors[k + checkpoint.VARIABLE_SUFFIX]) for k in wanted)
    listed = checkpoint.list_variables(prefix)
    assert listed["save_counter" + checkpoint.VARIABLE_SUFFIX] == ((), "int64")


def test_bundle_written_natively_is_read_by_the_oracle(tmp_path):
    rng = np.random.default_rng(1)
    w = _weights(rng, n=120)  # > 4 KB of index entries: several data blocks
    prefix = str(tmp_path / "ckpt" / "final.ckpt")
    checkpoint.save_variables(prefix, w)
    assert sorted(os.listdir(tmp_path / "ckpt")) == ["checkpoint", "final.ckpt.data-00000-of-00001", "final.ckpt.index"]
    assert 'model_checkpoint_path: "final.ckpt"' in open(tmp_path / "ckpt" / "checkpoint").read()
    got = bundle_oracle.read_bundle(prefix)
    assert set(got) == {k + checkpoint.VARIABLE_SUFFIX for k in w}
    for k, v in w.items():
        assert np.array_equal(got[k + checkpoint.VARIABLE_SUFFIX], v) and got[k + checkpoint.VARIABLE_SUFFIX].shape == v.shape
    table = bundle_oracle.read_table(prefix + ".index")
    assert [k for k, _ in table] == sorted(k for k, _ in table) and table[0][0] == b""
    assert len(open(prefix + ".data-00000-of-00001", "rb").read()) >= sum(v.nbytes for v in w.values())
    again = checkpoint.load_variables(prefix, {k: v.shape for k, v in w.items()})
    assert all(np.array_equal(again[k], w[k]) for k in w)


def test_object_graph_restore_follows_edges_not_key_spelling(tmp_path):
    """TensorFlow's object-based restore matches variables by walking named edges from the root; keys may be spelled differently
    (a Sequential reached through ``layer_with_weights-N``, shared objects keyed by their first path)."""
    rng = np.random.default_rng(2)
    k1, k2 = rng.standard_normal((3, 2)).astype(np.float32), rng.standard_normal((2,)).astype(np.float32)
    g = checkpoint.ObjectGraph()
    # root -(model)-> 1 -(mlp)-> 2 -(layer_with_weights-0)-> 3 -(kernel)-> 4, -(bias)-> 5 ; root -(optimizer)-> 6
    edges = [{"model": 1, "optimizer": 6}, {"mlp": 2}, {"layer_with_weights-0": 3, "layer-0": 3}, {"kernel": 4, "bias": 5}, {}, {}, {}]
    keys = {4: "some/other/spelling/kernel" + checkpoint.VARIABLE_SUFFIX, 5: "some/other/spelling/bias" + checkpoint.VARIABLE_SUFFIX}
    from collections import OrderedDict

    for i, e in enumerate(edges):
        g.children.append(OrderedDict(e))
        g.attributes.append(OrderedDict({"VARIABLE_VALUE": keys[i]} if i in keys else {}))
        g.full_names.append({})
    assert checkpoint.ObjectGraph.parse(g.serialize()).children == g.children
    lib = io_lib.load_library()
    prefix = str(tmp_path / "g.ckpt")
    w = io_lib.check_handle(lib.fdio_bundle_writer_create(prefix.encode()))
    for key, arr in ((keys[4], k1), (keys[5], k2)):
        dims = (ctypes.c_int64 * arr.ndim)(*arr.shape)
        io_lib.check(lib.fdio_bundle_writer_add(w, key.encode(), 1, arr.ndim, dims, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
    blob = g.serialize()
    io_lib.check(lib.fdio_bundle_writer_add(w, checkpoint.OBJECT_GRAPH_KEY.encode(), checkpoint.DT_STRING, 0, None, blob, len(blob)))
    io_lib.check(lib.fdio_bundle_writer_finish(w))
    got = checkpoint.load_variables(prefix, {"model/mlp/layer_with_weights-0/kernel": (3, 2), "model/mlp/layer_with_weights-0/bias": (2,)})
    assert np.array_equal(got["model/mlp/layer_with_weights-0/kernel"], k1) and np.array_equal(got["model/mlp/layer_with_weights-0/bias"], k2)
    with checkpoint.Bundle(prefix) as b:
        assert list(b.object_graph().variables()) == ["model/mlp/layer_with_weights-0/kernel", "model/mlp/layer_with_weights-0/bias"]
    with pytest.raises(KeyError, match="model/mlp/layer_with_weights-1/kernel"):
        checkpoint.load_variables(prefix, {"model/mlp/layer_with_weights-1/kernel": (3, 2)})
    with pytest.raises(ValueError, match="shape"):
        checkpoint.load_variables(prefix, {"model/mlp/layer_with_weights-0/kernel": (2, 3)})


def test_checkpoint_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(3)
    w = _weights(rng, n=3)
    prefix = str(tmp_path / "c.ckpt")
    checkpoint.save_variables(prefix, w)
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(io_lib.IOError_, match="checksum mismatch"):
        checkpoint.load_variables(prefix, {k: v.shape for k, v in w.items()})
    index = bytearray(open(prefix + ".index", "rb").read())
    index[5] ^= 1
    open(prefix + ".index", "wb").write(index)
    with pytest.raises(io_lib.IOError_, match="block checksum mismatch"):
        checkpoint.Bundle(prefix)
    open(prefix + ".index", "wb").write(index[:-1])
    with pytest.raises(io_lib.IOError_, match="magic"):
        checkpoint.Bundle(prefix)
    with pytest.raises(FileNotFoundError):
        checkpoint.Bundle(str(tmp_path / "missing.ckpt"))


# ------------------------------------------------------------------------------------------------------------------------ reference-run goldens
def _normalise(x):
    """Tuples -> lists, numpy / torch scalars -> Python, bytes -> str: the shape the golden JSON has."""
    if isinstance(x, dict):
        return {k: _normalise(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_normalise(v) for v in x]
    if isinstance(x, bytes):
        return x.decode("utf-8")
    if isinstance(x, np.ndarray):
        return _normalise(x.tolist())
    if isinstance(x, np.generic):
        return _normalise(x.item())
    return x


def _approx_equal(a, b, path=""):
    if isinstance(b, dict):
        assert isinstance(a, dict) and sorted(a) == sorted(b), path
        for k in b:
            _approx_equal(a[k], b[k], path + "/" + str(k))
    elif isinstance(b, list):
        assert isinstance(a, list) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _approx_equal(x, y, "%s[%d]" % (path, i))
    elif isinstance(b, float):
        assert a == pytest.approx(b, rel=1e-12, abs=1e-12), path
    else:
        assert a == b, (path, a, b)


@pytest.mark.parametrize("name", ["crello", "rico"])
def test_dataspec_matches_the_reference_run_golden(name):
    """``tests/golden/make_dataspec_golden.py`` ran the reference's own ``DataSpec`` (its YAML specs, ``_create_lookup``,
    ``make_input_columns``, ``parse_fn``, ``unbatch``) on the committed TFRecord fixture; this class must reproduce every output."""
    import json

    golden_dir = os.path.join(ROOT, "tests", "golden")
    meta = json.load(open(os.path.join(golden_dir, "dataspec_%s.json" % name)))
    arrays = np.load(os.path.join(golden_dir, "dataspec_%s.npz" % name))
    root = os.path.join(golden_dir, "dataspec_" + name)
    spec = DataSpec(name, root, batch_size=4)
    # the schema restated in BUILTIN_SPECS is the reference's YAML
    assert list(spec.columns.keys()) == meta["column_order"]
    _approx_equal(_normalise(json.loads(json.dumps(spec.columns))), meta["spec_columns"])
    _approx_equal(_normalise(spec.make_input_columns()), meta["input_columns"])
    assert spec.size("train") == meta["size"] and spec.steps_per_epoch("train") == meta["steps_per_epoch"]
    shard = TFRecordFile(os.path.join(root, "train-00000-of-00001.tfrecord"), verify_crc=2)
    records = [shard.record(i) for i in range(len(shard))]
    assert records == DO.read_tfrecord(shard.path)
    batch = spec.parse_fn(records)
    assert sorted(k for k in batch if not isinstance(batch[k], np.ndarray)) == sorted(arrays.files)
    for k in arrays.files:
        got = batch[k].numpy()
        assert str(got.dtype) == meta["dtypes"][k] and got.shape == arrays[k].shape and np.array_equal(got, arrays[k]), k
    for k, v in meta["strings"].items():
        assert _normalise(batch[k]) == v, k
    _approx_equal(_normalise(spec.unbatch(batch)), meta["unbatch"])
    # the same batch through the dataset iterator
    first = next(iter(spec.make_dataset("train", batch_size=len(records), shuffle=False)))
    for k in arrays.files:
        assert np.array_equal(first[k].numpy(), arrays[k]), k


def test_sharded_datasets_partition_the_documents(crello_dir):
    """``shard=(rank, world)``: the data-parallel ranks draw one seeded document order and split it document by document."""
    root, _ = crello_dir
    spec = DataSpec("crello", root, batch_size=4)
    ids = lambda ds, n=None: [bytes(x) for i, b in zip(range(n or 10 ** 9), ds) for x in b["id"][:, 0]]
    whole = ids(spec.make_dataset("train", shuffle=True, seed=6, strings=True))
    parts = [ids(spec.make_dataset("train", shuffle=True, seed=6, strings=True, shard=(r, 3))) for r in range(3)]
    # 37 documents over 3 ranks: every rank gets ceil(37 / 3) = 13 (all ranks must run the same number of batches: their metric rows are
    # all-reduced collectively); the last round wraps to the head of the pass, so the two short ranks repeat documents 1 and 2 of the order
    assert [len(p) for p in parts] == [13, 13, 13] and set(sum(parts, [])) == set(whole)
    assert all(parts[r][:12] == whole[r::3][:12] for r in range(3))
    assert parts[0][12] == whole[36] and parts[1][12] == whole[0] and parts[2][12] == whole[1]
    rep = [ids(spec.make_dataset("train", shuffle=True, seed=6, strings=True, repeat=True, shard=(r, 2)), n=20) for r in range(2)]
    assert len(rep[0]) == len(rep[1]) == 80 and not set(rep[0][:18]) & set(rep[1][:18])
    with pytest.raises(ValueError):
        spec.make_dataset("train", shard=(3, 3))


# ------------------------------------------------------------------------------------------------------------------------ property test
def _random_spec_and_records(rng):
    """A random column spec in the YAML grammar of the reference (dtype / shape / is_sequence / lookup / discretize / max) with a
    vocabulary, and random documents for it -- including empty documents, out-of-vocabulary tokens where the lookup has an OOV index,
    mask tokens, values on and beyond the discretiser's boundaries, negative integers, and unpacked repeated fields."""
    columns, vocabulary = {}, {}
    n_cols = int(rng.integers(1, 7))
    for ci in range(n_cols):
        name = "c%d" % ci
        dtype = ["int64", "float32", "string"][int(rng.integers(3))]
        col = {"dtype": dtype}
        if rng.random() < 0.7:
            col["is_sequence"] = True
        width = int(rng.integers(1, 5)) if rng.random() < 0.4 else 1
        if width > 1:
            col["shape"] = [width]
        kind = rng.random()
        if dtype == "string" or (dtype == "int64" and kind < 0.4):
            size = int(rng.integers(1, 9))
            mask = rng.random() < 0.5
            oov = int(rng.random() < 0.6)
            if dtype == "string":
                tokens = ["tok%d" % k for k in range(size)]
                col["lookup"] = {"num_oov_indices": oov, "mask_token": "" if mask else None}
                if not mask and not oov:
                    col["lookup"]["mask_token"] = ""  # padding "" must stay representable
            else:
                tokens = [int(x) for x in rng.choice(np.arange(-20, 40), size=size, replace=False)]
                tokens = [t for t in tokens if t != 0] or [7]
                col["lookup"] = {"num_oov_indices": oov, "mask_token": 0 if (mask or not oov) else None}
            if rng.random() < 0.5:
                col["min_freq"] = 5
                vocabulary[name] = {str(t): int(rng.integers(1, 10)) for t in tokens}
                if all(f < 5 for f in vocabulary[name].values()):
                    vocabulary[name][str(tokens[0])] = 9
            else:
                vocabulary[name] = tokens
        elif dtype != "string" and kind < 0.75:
            lo = float(rng.integers(-3, 1))
            col["discretize"] = {"min": lo, "max": lo + float(rng.integers(1, 300)), "bins": int(rng.integers(2, 70))}
        elif dtype == "int64":
            col["max"] = 9
        columns[name] = col
    pre = DO.make_preprocessors(columns, vocabulary)
    records = []
    for _ in range(int(rng.integers(1, 6))):
        steps = int(rng.integers(0, 6))
        context, lists = {}, {}
        for name, col in columns.items():
            width = int(np.prod(col.get("shape", (1,))))

            def draw():
                lk = pre.get(name) if "lookup" in col else None
                if lk is not None:
                    first = (0 if lk.mask is None else 1) + lk.num_oov
                    pool = list(lk.tokens[first:]) + ([lk.mask] if lk.mask is not None else []) + (["zz unseen" if lk.is_string else 12345] if lk.num_oov else [])
                    return [pool[int(rng.integers(len(pool)))] for _ in range(width)]
                if col["dtype"] == "string":
                    return ["s%d" % rng.integers(100) for _ in range(width)]
                if col["dtype"] == "float32":
                    d = col.get("discretize")
                    if d and rng.random() < 0.4:  # exactly on a boundary, or outside the range
                        edges = list(np.linspace(d["min"], d["max"], d["bins"])) + [d["min"] - 1.0, d["max"] + 1.0]
                        return [float(edges[int(rng.integers(len(edges)))]) for _ in range(width)]
                    return [float(np.float32(rng.normal() * 50)) for _ in range(width)]
                return [int(rng.integers(-5, 300)) for _ in range(width)]

            def feature(values):
                if col["dtype"] == "string" or rng.random() < 0.7:
                    return encode_feature(values, col["dtype"])
                if col["dtype"] == "float32":  # unpacked: one fixed32 field per value
                    return _ld(2, b"".join(b"\x0d" + struct.pack("<f", v) for v in values))
                return _ld(3, b"".join(b"\x08" + _varint(v & ((1 << 64) - 1)) for v in values))

            if col.get("is_sequence"):
                lists[name] = [feature(draw()) for _ in range(steps)]
            else:
                context[name] = feature(draw())
        records.append(encode_sequence_example(context, lists))
    return columns, vocabulary, records


def test_random_specs_and_documents_match_the_oracle(tmp_path):
    """Property test over random schemas and documents: the native parser and the pure-Python restatement agree bit for bit."""
    rng = np.random.default_rng(2024)
    checked = 0
    for trial in range(60):
        columns, vocabulary, records = _random_spec_and_records(rng)
        d = tmp_path / ("t%d" % trial)
        d.mkdir()
        spec = _tiny_spec(d, columns, vocabulary)
        want = DO.parse_fn(columns, vocabulary, records)
        got = spec.parse_fn(records)
        _assert_batches_equal(got, want)
        # the input columns the model is sized from follow the same vocabulary arithmetic
        cols = spec.make_input_columns()
        pre = DO.make_preprocessors(columns, vocabulary)
        for name, col in columns.items():
            if "lookup" in col:
                assert cols[name]["input_dim"] == len(pre[name].tokens), name
            elif "discretize" in col:
                assert cols[name]["input_dim"] == col["discretize"]["bins"], name
        checked += sum(v.size for v in want.values())
    assert checked > 500


def test_reference_style_checkpoint_with_optimizer_slots_restores_the_model_variables(tmp_path):
    """A bundle shaped like the reference's ``best.ckpt`` (Keras ``save_weights`` of the compiled model: every model variable of SURVEY.md
    Appendix B, Adam's ``iter`` / hyper-parameters / ``m`` and ``v`` slots, metric variables, no object graph) written by the independent
    oracle writer: ``load_variables`` picks exactly the model variables, for every architecture switch that adds variables."""
    from oracle import mfp_oracle as O

    rng = np.random.default_rng(8)
    for kwargs in ({}, {"input_dtype": "shuffled_set"}, {"context": "id"}, {"context": "length"}):
        specs = O.variable_specs(make_input_columns("rico"), 2, 8, kwargs.get("input_dtype", "set"), kwargs.get("context"))  # small D: names matter here
        weights = {name: rng.standard_normal(shape).astype(np.float32) for name, (shape, _, _) in specs.items()}
        tensors = {name + checkpoint.VARIABLE_SUFFIX: w for name, w in weights.items()}
        for name, w in weights.items():  # Adam slots sit next to their variables
            for slot in ("m", "v"):
                tensors["%s/.OPTIMIZER_SLOT/optimizer/%s%s" % (name, slot, checkpoint.VARIABLE_SUFFIX)] = np.zeros_like(w)
        tensors["optimizer/iter" + checkpoint.VARIABLE_SUFFIX] = np.asarray(1234, dtype=np.int64)
        for hp in ("beta_1", "beta_2", "decay", "learning_rate"):
            tensors["optimizer/%s%s" % (hp, checkpoint.VARIABLE_SUFFIX)] = np.asarray(0.5, dtype=np.float32)
        tensors["keras_api/metrics/0/total" + checkpoint.VARIABLE_SUFFIX] = np.asarray(1.0, dtype=np.float32)
        prefix = str(tmp_path / ("best_%s.ckpt" % "_".join(map(str, kwargs.values()))))
        bundle_oracle.write_bundle(prefix, tensors, entries_per_block=7, restart_interval=4)
        got = checkpoint.load_variables(prefix, {k: v.shape for k, v in weights.items()})
        assert list(got) == list(weights)
        for k in weights:
            assert np.array_equal(got[k], weights[k]), k
    # keys spelled with an extra wrapper edge are still found when the match is unique and the shape agrees ...
    tensors = {"root/model/decoder/decoders/left/kernel" + checkpoint.VARIABLE_SUFFIX: np.ones((8, 64), np.float32)}
    prefix = str(tmp_path / "wrapped.ckpt")
    bundle_oracle.write_bundle(prefix, tensors)
    got = checkpoint.load_variables(prefix, {"model/decoder/decoders/left/kernel": (8, 64)})
    assert got["model/decoder/decoders/left/kernel"].shape == (8, 64)
    # ... and refused when it is ambiguous
    tensors["other/model/decoder/decoders/left/kernel" + checkpoint.VARIABLE_SUFFIX] = np.zeros((8, 64), np.float32)
    bundle_oracle.write_bundle(prefix, tensors)
    with pytest.raises(KeyError):
        checkpoint.load_variables(prefix, {"model/decoder/decoders/left/kernel": (8, 64)})
