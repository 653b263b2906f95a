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
