"""GPU tests of the widened rows either side of the step: ``train.py``'s own flow -- ``DataSpec`` over
TFRecord shards -> ``make_input_columns`` -> ``MFP`` -> ``fit`` with the dataset -> TensorFlow-format ``save_weights`` / ``load_weights`` --
and bit-equality of a step fed from files with the same step fed from the arrays that were exported."""
import os

import numpy as np
import pytest

from flex_dm_b200.dataspec import DataSpec
from flex_dm_b200.synthetic import write_synthetic_dataset

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,method", [("rico", "elem_pos_attr"), ("crello", "random")])
def test_train_py_flow_from_tfrecords(tmp_path, name, method):
    import torch

    from flex_dm_b200 import checkpoint
    from flex_dm_b200.mfp import MFP, Adam

    root = str(tmp_path / "data")
    written = write_synthetic_dataset(root, name, {"train": 24, "val": 8}, seq_len=10, shards=2, seed=5)
    dataspec = DataSpec(name, root, batch_size=8)  # train.py:38-43
    input_columns = dataspec.make_input_columns()
    model = MFP(input_columns, num_blocks=1, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=3)
    model.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0), run_eagerly=True)

    # a batch parsed from the shards is the batch that was exported: the step's metrics are identical, bit for bit
    from_files = next(iter(dataspec.make_dataset("train", shuffle=False)))
    direct = {k: v[:8] for k, v in written["train"][0].items()}
    S = from_files["left"].shape[1]
    direct = {k: (v[:, :S] if v.ndim == 3 else v) for k, v in direct.items()}
    model._step = 500
    a = model.test_step(from_files).cpu().numpy()
    model._step = 500
    b = model.test_step(direct).cpu().numpy()
    assert np.array_equal(a, b)
    if torch.cuda.is_available():
        assert all(t.is_pinned() for t in from_files.values())  # parsed straight into pinned memory

    train = dataspec.make_dataset("train", shuffle=True, repeat=True, cache=True)  # train.py:44-46
    val = dataspec.make_dataset("val", cache=True)  # train.py:47-48
    history = model.fit(train, steps_per_epoch=dataspec.steps_per_epoch("train"), epochs=4, validation_data=val,
                        validation_steps=dataspec.steps_per_epoch("val"), verbose=0)
    assert len(history) == 4 and history[-1]["loss"] < history[0]["loss"] and np.isfinite(history[-1]["val_loss"])

    path = os.path.join(str(tmp_path), "checkpoints", "final.ckpt")  # train.py:94-97
    model.save_weights(path)
    assert checkpoint.is_tf_checkpoint(path)
    listed = checkpoint.list_variables(path)
    assert "model/encoder/input_layer/left/embeddings/.ATTRIBUTES/VARIABLE_VALUE" in listed and checkpoint.OBJECT_GRAPH_KEY in listed
    other = MFP(input_columns, num_blocks=1, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=77)
    other.compile(optimizer="adam")  # eval.py:170
    other.load_weights(path)  # eval.py:172
    for k, v in model.get_weights().items():
        assert np.array_equal(v, other.get_weights()[k]), k
    model._step = other._step = 900
    other.seed = model.seed
    assert other.evaluate(dataspec.make_dataset("val")) == pytest.approx(model.evaluate(dataspec.make_dataset("val")), rel=1e-6)
    # a model of another shape refuses the checkpoint like Keras does
    wrong = MFP(input_columns, num_blocks=2, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=1)
    with pytest.raises(KeyError, match="seq2seq_1"):
        wrong.load_weights(path)


@pytest.mark.parametrize("name", ["crello", "rico"])
def test_device_cached_dataset_matches_streaming(tmp_path, name):
    """``make_dataset(cache="device")`` keeps the parsed split ragged in HBM and cuts batches out of it with ``mfp_gather_documents``:
    every batch must be bit-identical to the one the host parser streams for the same seed (ragged documents, a partial last batch,
    batch-maximum and fixed padding), and training from it must behave the same."""
    import torch

    from flex_dm_b200.mfp import MFP, Adam

    root = str(tmp_path / "data")
    write_synthetic_dataset(root, name, {"train": 53, "val": 7}, seq_len=13, shards=3, seed=9)
    dataspec = DataSpec(name, root, batch_size=8)
    for kwargs in ({"shuffle": True, "seed": 3}, {"shuffle": False, "pad_to": 16}, {"shuffle": 7, "seed": 1, "batch_size": 5}):
        streamed = list(dataspec.make_dataset("train", prefetch=0, **kwargs))
        cached_ds = dataspec.make_dataset("train", cache="device", **kwargs)
        cached = list(cached_ds)
        assert len(cached) == len(streamed) == -(-53 // kwargs.get("batch_size", 8))
        for a, b in zip(cached, streamed):
            assert list(a.keys()) == list(b.keys())
            for k in b:
                assert a[k].is_cuda and a[k].dtype == b[k].dtype and tuple(a[k].shape) == tuple(b[k].shape), k
                assert torch.equal(a[k].cpu(), b[k]), k
        assert cached_ds.nbytes() > 0 and len(cached_ds) == 53
    # a second pass reshuffles like the streaming dataset does (same epoch counter semantics)
    s_ds = dataspec.make_dataset("train", prefetch=0, shuffle=True, seed=5)
    c_ds = dataspec.make_dataset("train", cache="device", shuffle=True, seed=5)
    for _ in range(2):
        for a, b in zip(c_ds, s_ds):
            assert torch.equal(a["left"].cpu(), b["left"])
    # train.py's loop on the cached dataset: no host parsing and no H2D copies of the columns in the steady state
    model = MFP(dataspec.make_input_columns(), num_blocks=1, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=2)
    model.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    model.set_deterministic(True)  # fixed-order gradient reductions: the two runs below must agree bit for bit
    train = dataspec.make_dataset("train", shuffle=True, repeat=True, cache="device", seed=4)
    history = model.fit(train, steps_per_epoch=dataspec.steps_per_epoch("train"), epochs=3, validation_data=dataspec.make_dataset("val", cache="device"),
                        validation_steps=1, verbose=0)
    assert history[-1]["loss"] < history[0]["loss"]
    # ... and it is the same training run as from the streamed batches
    other = MFP(dataspec.make_input_columns(), num_blocks=1, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=2)
    other.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    other.set_deterministic(True)
    h2 = other.fit(dataspec.make_dataset("train", shuffle=True, repeat=True, seed=4), steps_per_epoch=dataspec.steps_per_epoch("train"), epochs=3,
                   validation_data=dataspec.make_dataset("val"), validation_steps=1, verbose=0)  # (validation advances the RNG step counter too)
    assert [h["loss"] for h in h2] == [h["loss"] for h in history]  # deterministic mode: identical batches -> identical runs
    for name, w in model.get_weights().items():
        assert np.array_equal(other.get_weights()[name], w), name
