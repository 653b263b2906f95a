"""train.py:79-88 with the reference's callback list on the GPU engine (the callback protocol itself is covered on the CPU by
tests/test_api_surface.py; this is the end-to-end flow that leaves ``best.ckpt`` for eval.py:169-172)."""
import itertools
import os
from types import SimpleNamespace

import numpy as np
import pytest

from flex_dm_b200.spec import make_input_columns, make_synthetic_batch

pytestmark = pytest.mark.gpu


def test_fit_with_the_reference_callback_list_leaves_the_best_checkpoint(tmp_path):
    from flex_dm_b200.callbacks import get_callbacks
    from flex_dm_b200.mfp import MFP, Adam

    cols = make_input_columns("rico")
    model = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", latent_dim=256, dropout=0.1, l2=1e-2, seed=1)
    model.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0), run_eagerly=True)
    batches = [make_synthetic_batch(cols, 8, 12, seed=i, lengths="ragged") for i in range(3)]
    args = SimpleNamespace(job_dir=str(tmp_path))
    ckpt = os.path.join(args.job_dir, "checkpoints", "best.ckpt")
    snapshots = {}
    callbacks = get_callbacks(args, None, ckpt) + [lambda epoch, logs, m: snapshots.__setitem__(epoch, m.get_weights())]
    history = model.fit(itertools.cycle(batches), steps_per_epoch=3, epochs=4, validation_data=batches[:1], validation_steps=1, validation_freq=2,
                        callbacks=callbacks, verbose=0)
    assert len(history) == 4 and [("val_total_score" in h) for h in history] == [False, True, False, True]
    best_epoch, best = None, -np.inf  # ModelCheckpoint(monitor="val_total_score", mode="max", save_best_only=True): strict improvement
    for epoch, logs in enumerate(history):
        if "val_total_score" in logs and logs["val_total_score"] > best:
            best_epoch, best = epoch, logs["val_total_score"]
    assert best_epoch is not None and callbacks[1].best == best
    assert os.path.exists(ckpt + ".index") and os.path.exists(ckpt + ".data-00000-of-00001")
    other = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", latent_dim=256, dropout=0.1, l2=1e-2, seed=99)
    other.compile(optimizer="adam")
    other.load_weights(ckpt)  # eval.py:169-172
    for name, w in snapshots[best_epoch].items():
        assert np.array_equal(other.get_weights()[name], w), name
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator

    events = EventAccumulator(os.path.join(args.job_dir, "logs", "validation"))
    events.Reload()
    got = [(e.step, e.value) for e in events.Scalars("epoch_total_score")]
    assert [s for s, _ in got] == [1, 3] and got[0][1] == pytest.approx(history[1]["val_total_score"], rel=1e-6)


def test_train_function_end_to_end(tmp_path):
    """``training.train(args)`` -- everything ``python -m mfp`` does after parsing (train.py:16-97) -- on a small synthetic rico export: the
    job directory holds args.json, TensorBoard logs, best.ckpt and final.ckpt; final.ckpt reproduces the returned test metrics."""
    from types import SimpleNamespace

    from flex_dm_b200 import checkpoint, training
    from flex_dm_b200.dataspec import DataSpec
    from flex_dm_b200.mfp import MFP
    from flex_dm_b200.synthetic import write_synthetic_dataset

    data = str(tmp_path / "data")
    write_synthetic_dataset(data, "rico", {"train": 24, "val": 8, "test": 8}, seq_len=10, shards=2, seed=5)
    args = SimpleNamespace(dataset_name="rico", data_dir=data, job_dir=str(tmp_path / "job"), batch_size=8, weights=None, latent_dim=256, num_blocks=1,
                           arch_type="oneshot", block_type="deepsvg", l2=1e-2, dropout=0.1, masking_method="elem_pos_attr", seq_type="default",
                           context=None, input_dtype="set", learning_rate=1e-3, num_epochs=3, validation_freq=1, verbose=0, seed=2)
    results = training.train(args)
    assert set(results) >= {"loss", "total_score", "left_score"} and all(np.isfinite(v) for v in results.values())
    ckpts = os.path.join(args.job_dir, "checkpoints")
    assert os.path.exists(os.path.join(args.job_dir, "args.json")) and os.path.isdir(os.path.join(args.job_dir, "logs", "validation"))
    assert checkpoint.is_tf_checkpoint(os.path.join(ckpts, "best.ckpt")) and checkpoint.is_tf_checkpoint(os.path.join(ckpts, "final.ckpt"))
    spec = DataSpec("rico", data, batch_size=8)
    other = MFP(spec.make_input_columns(), num_blocks=1, masking_method="elem_pos_attr", latent_dim=256, dropout=0.1, l2=1e-2, seed=2)
    other.compile(optimizer="adam")
    other.load_weights(os.path.join(ckpts, "final.ckpt"))
    again = dict(zip(other.metrics_names, other.evaluate(spec.make_dataset("test"))))
    # same weights, same Philox key; the corruption streams are indexed by the step counter, which differs between the two runs, so
    # the masked positions differ: scores agree statistically, the L2 part of the loss exactly
    assert 0.0 <= again["total_score"] <= 1.0 and again["total_score"] == pytest.approx(results["total_score"], abs=0.3)
