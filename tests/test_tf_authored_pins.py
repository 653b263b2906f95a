"""Pins against TensorFlow-AUTHORED code that ships inside the ``tensorboard`` package (no TensorFlow needed): its pure-Python TFRecord
writer / reader with masked CRC-32C (``tensorboard.summary.writer.record_writer``, ``tensorboard.compat.tensorflow_stub.pywrap_tensorflow``,
"packing defined in tensorflow") and its generated protobuf classes (``tensorboard.compat.proto``: ``TrackableObjectGraph`` -- what a Keras
checkpoint stores under ``_CHECKPOINTABLE_OBJECT_GRAPH`` --, ``DataType``, ``TensorShapeProto``).  They turn three of the "restated from the
published description" items of DESIGN.md section 2 into checks against the TensorFlow team's own implementation: record framing and
CRC masking, the object-graph message, and the dtype numbering of bundle entries.  (``tf.train.SequenceExample`` and
``BundleEntryProto`` are not among the protos tensorboard ships; those stay pinned by ``google.protobuf`` encoders and the oracles.)"""
import os

import numpy as np
import pytest

tensorboard = pytest.importorskip("tensorboard")

from flex_dm_b200 import checkpoint, io_lib  # noqa: E402
from flex_dm_b200.dataspec import TFRecordFile, write_tfrecord  # noqa: E402


def test_crc32c_and_masking_equal_tensorboards():
    from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tb

    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 8, 9, 63, 64, 65, 4096, 100003):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert io_lib.crc32c(data) == tb.crc32c(data), n
        assert io_lib.masked_crc32c(data) == tb.masked_crc32c(data), n


def test_records_written_by_tensorboard_are_read_natively(tmp_path):
    """TensorBoard's RecordWriter -> libflexdm_io's mmap reader with length and payload CRC verification (verify_crc=2)."""
    from tensorboard.summary.writer.record_writer import RecordWriter

    rng = np.random.default_rng(1)
    records = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (0, 1, 5, 4096, 70000)]
    path = str(tmp_path / "theirs.tfrecord")
    with open(path, "wb") as f:
        w = RecordWriter(f)
        for r in records:
            w.write(r)
        w.flush()
    got = TFRecordFile(path, verify_crc=2)
    assert [got.record(i) for i in range(len(got))] == records


def test_records_written_natively_are_read_by_tensorboard(tmp_path):
    from tensorboard.compat.tensorflow_stub import errors
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import PyRecordReader_New

    rng = np.random.default_rng(2)
    records = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1, 7, 300, 65536)]
    path = str(tmp_path / "ours.tfrecord")
    write_tfrecord(path, records)
    reader = PyRecordReader_New(path)
    got = []
    while True:
        try:
            reader.GetNext()  # verifies both CRCs, raises DataLossError on a mismatch
        except errors.OutOfRangeError:
            break
        got.append(reader.record())
    assert got == records


def test_event_files_are_tfrecords_of_event_protos(tmp_path):
    """An event file written by the TensorBoard callback (flex_dm_b200/callbacks.py) is itself a TFRecord stream: the native reader
    indexes it and the records decode as ``Event`` messages holding the scalars."""
    from tensorboard.compat.proto import event_pb2

    from flex_dm_b200.callbacks import TensorBoard

    cb = TensorBoard(str(tmp_path))
    cb.on_epoch_end(0, {"loss": 3.5})
    cb.on_epoch_end(1, {"loss": 2.25})
    cb.on_train_end()
    (name,) = os.listdir(tmp_path / "train")
    f = TFRecordFile(str(tmp_path / "train" / name), verify_crc=2)
    events = [event_pb2.Event.FromString(f.record(i)) for i in range(len(f))]
    scalars = [(e.step, v.tag, v.tensor.float_val[0] if v.HasField("tensor") else v.simple_value) for e in events for v in e.summary.value]
    assert scalars == [(0, "epoch_loss", 3.5), (1, "epoch_loss", 2.25)]


NAMES = ["model/encoder/input_layer/left/embeddings", "model/blocks/seq2seq/seq2seq_0/attn/dense_query/kernel",
         "model/blocks/seq2seq/seq2seq_0/attn/dense_query/bias", "model/decoder/decoders/left/kernel"]


def test_object_graph_written_here_parses_as_tensorflows_message():
    from tensorboard.compat.proto.trackable_object_graph_pb2 import TrackableObjectGraph

    ours = checkpoint.ObjectGraph.from_variable_paths(NAMES)
    theirs = TrackableObjectGraph.FromString(ours.serialize())
    assert len(theirs.nodes) == len(ours.children)
    for node, children, attributes in zip(theirs.nodes, ours.children, ours.attributes):
        assert [(c.local_name, c.node_id) for c in node.children] == list(children.items())
        assert [(a.name, a.checkpoint_key) for a in node.attributes] == list(attributes.items())
    # following TensorFlow's message edge by edge from the root reaches every variable under the key save_weights stores it at
    for name in NAMES:
        node = theirs.nodes[0]
        for edge in name.split("/"):
            (ref,) = [c for c in node.children if c.local_name == edge]
            node = theirs.nodes[ref.node_id]
        (attr,) = node.attributes
        assert attr.name == "VARIABLE_VALUE" and attr.checkpoint_key == name + checkpoint.VARIABLE_SUFFIX
    assert theirs.SerializeToString(deterministic=True) == ours.serialize()  # canonical field order, no stray bytes


def test_object_graph_built_with_tensorflows_message_is_walked_here():
    """A Keras-style graph assembled with the TensorFlow-authored classes: wrapper edges, an optimizer with slot variables, keys spelled
    differently from the attribute paths -- the restore logic finds the model's variables by edges and ignores the rest."""
    from tensorboard.compat.proto.trackable_object_graph_pb2 import TrackableObjectGraph

    g = TrackableObjectGraph()
    root = g.nodes.add()
    index = {(): 0}
    for name in NAMES:
        parts = tuple(name.split("/"))
        for depth in range(1, len(parts) + 1):
            prefix = parts[:depth]
            if prefix not in index:
                index[prefix] = len(g.nodes)
                g.nodes.add()
                g.nodes[index[prefix[:-1]]].children.add(node_id=index[prefix], local_name=prefix[-1])
        g.nodes[index[parts]].attributes.add(name="VARIABLE_VALUE", full_name="v%d" % index[parts], checkpoint_key="weights/%d" % index[parts])
    opt = len(g.nodes)
    g.nodes.add()
    root.children.add(node_id=opt, local_name="optimizer")
    slot = len(g.nodes)
    g.nodes.add().attributes.add(name="VARIABLE_VALUE", checkpoint_key="optimizer/m/0")
    g.nodes[opt].slot_variables.add(original_variable_node_id=index[tuple(NAMES[0].split("/"))], slot_name="m", slot_variable_node_id=slot)
    g.nodes[opt].has_checkpoint_values.value = True
    ours = checkpoint.ObjectGraph.parse(g.SerializeToString())
    for name in NAMES:
        assert ours.variable_key(name.split("/")) == "weights/%d" % index[tuple(name.split("/"))]
    assert ours.variable_key(["optimizer"]) is None and ours.walk(["model", "nothing"]) is None
    assert set(ours.variables()) == set(NAMES)  # the slot variable hangs on no child edge: not a model variable


def test_bundle_dtype_numbers_are_tensorflows():
    from tensorboard.compat.proto import types_pb2

    by_number = {v.number: n for n, v in types_pb2.DataType.DESCRIPTOR.values_by_name.items()}
    want = {np.float32: "DT_FLOAT", np.float64: "DT_DOUBLE", np.int32: "DT_INT32", np.uint8: "DT_UINT8", np.int16: "DT_INT16", np.int8: "DT_INT8",
            np.int64: "DT_INT64", np.bool_: "DT_BOOL", np.uint16: "DT_UINT16", np.float16: "DT_HALF", np.uint32: "DT_UINT32", np.uint64: "DT_UINT64"}
    assert {k: by_number[n] for n, k in ((n, checkpoint._DTYPES[n]) for n in checkpoint._DTYPES)} == want
    assert by_number[checkpoint.DT_STRING] == "DT_STRING"
