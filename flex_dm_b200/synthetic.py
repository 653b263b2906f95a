"""Synthetic datasets in the reference's on-disk layout (``root/count.json``, ``root/vocabulary.json``, ``root/<split>-*.tfrecord``;
``src/mfp/mfp/data/spec.py:29-36``): the real crello / rico records are external downloads, so benchmarks and tests of the input side
export ``make_synthetic_batch`` documents as TFRecords of ``tf.train.SequenceExample`` and read them back through
``DataSpec``.  Raw values are chosen so that the spec's lookup / discretisation maps them back to the batch's indices, and the
vocabularies so that ``DataSpec.make_input_columns()`` equals ``spec.make_input_columns(name)``."""
import copy
import json
import os
from collections import OrderedDict
from typing import Dict, List

import numpy as np

from .dataspec import BUILTIN_SPECS, encode_feature, encode_sequence_example, write_tfrecord
from .spec import CRELLO_TYPES, make_input_columns, make_synthetic_batch


def spec_for(name: str, max_length: int = 50) -> Dict:
    """The built-in column spec; ``max_length`` > 50 widens the ``length`` vocabulary (synthetic seq_len = 128 documents)."""
    spec = copy.deepcopy(BUILTIN_SPECS[name])
    if max_length > 50:
        spec["columns"]["length"]["lookup"]["vocabulary"]["max"] = int(max_length)
    return spec


def synthetic_vocabulary(name: str, input_columns: Dict) -> Dict:
    """vocabulary.json: token -> frequency; rare tokens sit below the spec's ``min_freq`` and must be filtered out by the reader."""
    columns = BUILTIN_SPECS[name]["columns"]
    vocab = OrderedDict()
    for key, column in columns.items():
        lookup = column.get("lookup")
        if lookup is None or (isinstance(lookup, dict) and "vocabulary" in lookup):
            continue
        n = input_columns[key]["input_dim"] - (1 if lookup.get("mask_token") is not None else 0) - lookup.get("num_oov_indices", 1)
        if key == "type" and name == "crello":
            tokens = CRELLO_TYPES[1:]
        elif column["dtype"] == "int64":
            tokens = [str(100 * (i + 1)) for i in range(n)]
        else:
            tokens = ["%s_%d" % (key, i) for i in range(n)]
        assert len(tokens) == n, (key, n)
        freq = OrderedDict((t, 1000 + i) for i, t in enumerate(tokens))
        if "min_freq" in column:
            freq["%s_rare" % key] = column["min_freq"] - 1
        vocab[key] = freq
    return vocab


def _raw_columns(columns: Dict, batch: Dict[str, np.ndarray], vocabulary: Dict) -> Dict[str, np.ndarray]:
    """Index batch -> the raw values a record stores."""
    raw = {}
    for key, column in columns.items():
        if column.get("demo_only"):
            continue
        x = batch[key]
        lookup = column.get("lookup")
        if lookup is not None:
            if isinstance(lookup, dict) and "vocabulary" in lookup:
                tokens = list(range(lookup["vocabulary"]["min"], lookup["vocabulary"]["max"] + 1))
                head = []
            else:
                tokens = [int(t) if column["dtype"] == "int64" else t for t, f in vocabulary[key].items() if f >= column.get("min_freq", 1)]
                mask = lookup.get("mask_token")
                head = ([] if mask is None else [mask]) + ["%s_unseen" % key] * lookup.get("num_oov_indices", 1)
            table = np.asarray(head + tokens, dtype=object)
            raw[key] = table[x]
        elif "discretize" in column:
            d = column["discretize"]
            scale = (d["max"] - d["min"]) / (d["bins"] - 1.0)
            centre = scale * (x + 0.5) + d["min"]
            raw[key] = np.minimum(np.floor(centre), d["max"]).astype(np.int64) if column["dtype"] == "int64" else centre.astype(np.float32)
        else:
            raw[key] = x
    return raw


def encode_documents(columns: Dict, batch: Dict[str, np.ndarray], vocabulary: Dict, id_prefix: str = "doc") -> List[bytes]:
    """One serialized SequenceExample per document of ``batch``; only the document's own elements are stored (no padding)."""
    raw = _raw_columns(columns, batch, vocabulary)
    B = batch["length"].shape[0]
    n_elem = batch["length"].reshape(B) + 1
    records = []
    for b in range(B):
        context, lists = OrderedDict(), OrderedDict()
        for key, column in columns.items():
            dtype = column["dtype"]
            if column.get("is_sequence"):
                if column.get("demo_only"):
                    lists[key] = [encode_feature(["%s-%d-%d" % (id_prefix, b, t)], dtype) for t in range(n_elem[b])]
                else:
                    lists[key] = [encode_feature(np.asarray(raw[key][b, t]).reshape(-1).tolist(), dtype) for t in range(n_elem[b])]
            elif column.get("demo_only"):
                context[key] = encode_feature(["%s-%d" % (id_prefix, b)], dtype)
            else:
                context[key] = encode_feature(np.asarray(raw[key][b]).reshape(-1).tolist(), dtype)
        records.append(encode_sequence_example(context, lists))
    return records


def write_synthetic_dataset(root: str, name: str, splits: Dict[str, int], seq_len: int = 50, lengths: str = "ragged", shards: int = 2,
                            seed: int = 0, chunk: int = 256) -> Dict[str, List[Dict[str, np.ndarray]]]:
    """Writes the dataset directory and returns, per split, the index batches that went in (document order = shard-major).
    With ``seq_len`` > 50 a ``<name>-spec.yml`` with the widened ``length`` vocabulary is written too: pass its path as ``DataSpec`` name."""
    os.makedirs(root, exist_ok=True)
    input_columns = make_input_columns(name, max_length=max(50, seq_len))
    spec = spec_for(name, max(50, seq_len))
    if seq_len > 50:
        import yaml

        with open(os.path.join(root, "%s-spec.yml" % name), "w") as f:
            yaml.safe_dump(json.loads(json.dumps(spec)), f, sort_keys=False)
    vocabulary = synthetic_vocabulary(name, input_columns)
    with open(os.path.join(root, "vocabulary.json"), "w") as f:
        json.dump(vocabulary, f)
    with open(os.path.join(root, "count.json"), "w") as f:
        json.dump(splits, f)
    written = {}
    for si, (split, count) in enumerate(splits.items()):
        per_shard = [count // shards + (1 if k < count % shards else 0) for k in range(shards)]
        written[split] = []
        for k, n in enumerate(per_shard):
            records: List[bytes] = []
            done = 0
            while done < n:
                b = min(chunk, n - done)
                batch = make_synthetic_batch(input_columns, b, seq_len, seed=seed + 7919 * si + 104729 * k + done, lengths=lengths)
                records += encode_documents(spec["columns"], batch, vocabulary, id_prefix="%s-%d-%d" % (split, k, done))
                written[split].append(batch)
                done += b
            write_tfrecord(os.path.join(root, "%s-%05d-of-%05d.tfrecord" % (split, k, shards)), records)
    return written
