#!/usr/bin/env python
"""Golden vectors of the input side, made by running the REFERENCE'S OWN ``DataSpec`` (``/root/reference/src/mfp/mfp/data/spec.py`` with
its ``crello-spec.yml`` / ``rico-spec.yml``) on small TFRecord fixtures, on top of the TensorFlow stand-in (``oracle/tf_standin``,
``tensorflow/data_io.py``).  Build container only (needs /root/reference); the fixtures and outputs are committed:

    tests/golden/dataspec_<name>/{count.json, vocabulary.json, train-00000-of-00001.tfrecord}     the dataset directory
    tests/golden/dataspec_<name>.json     make_input_columns(), size / steps_per_epoch, unbatch(parse_fn(records))
    tests/golden/dataspec_<name>.npz      parse_fn(records): every non-string column

Pinned by this: the column schemas (reference YAML vs ``flex_dm_b200.dataspec.BUILTIN_SPECS``), vocabulary handling, input_dim /
primary_label / loss_condition, parse order and casts, unbatch.  Not pinned: the TF ops underneath (record decoding, lookup and
bucketize semantics), which the stand-in restates -- see oracle/dataspec_oracle.py.
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
sys.path.insert(1, "/root/reference/src/mfp")
sys.path.insert(2, ROOT)

import tensorflow as tf  # noqa: E402,F401  (the stand-in)
from mfp.data.spec import DataSpec as RefDataSpec  # noqa: E402  (the reference)

from flex_dm_b200.synthetic import write_synthetic_dataset  # noqa: E402
from oracle import dataspec_oracle as DO  # noqa: E402


def plain(x):
    """JSON-able copy of the reference's return values (tensors, numpy scalars, bytes)."""
    if isinstance(x, dict):
        return {k: plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    if isinstance(x, bytes):
        return x.decode("utf-8")
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    if isinstance(x, np.ndarray):
        return plain(x.tolist())
    if isinstance(x, np.generic):
        return plain(x.item())
    return x


def main():
    for name, docs, seq_len in (("crello", 5, 6), ("rico", 7, 9)):
        root = os.path.join(HERE, "dataspec_" + name)
        shutil.rmtree(root, ignore_errors=True)
        write_synthetic_dataset(root, name, {"train": docs}, seq_len=seq_len, lengths="ragged", shards=1, seed=17)
        ref = RefDataSpec(name, root, batch_size=4)
        records = DO.read_tfrecord(os.path.join(root, "train-00000-of-00001.tfrecord"))
        columns = ref.make_input_columns()
        batch = ref.parse_fn(records)
        arrays = {k: v.detach().cpu().numpy() for k, v in batch.items() if hasattr(v, "detach")}
        strings = {k: plain(np.asarray(v)) for k, v in batch.items() if not hasattr(v, "detach")}
        items = ref.unbatch(dict(batch))
        meta = {"input_columns": plain(columns), "column_order": list(ref.columns.keys()), "size": ref.size("train"),
                "steps_per_epoch": ref.steps_per_epoch("train"), "strings": strings, "unbatch": plain(items),
                "dtypes": {k: str(v.dtype) for k, v in arrays.items()}, "spec_columns": plain(ref.columns)}
        with open(os.path.join(HERE, "dataspec_%s.json" % name), "w") as f:
            json.dump(meta, f, indent=1, sort_keys=True)
        np.savez_compressed(os.path.join(HERE, "dataspec_%s.npz" % name), **arrays)
        print(name, "columns:", len(columns), "batch:", {k: tuple(v.shape) for k, v in arrays.items()})


if __name__ == "__main__":
    main()
