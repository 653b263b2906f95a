"""train.py:79-88 with the reference's callback list on the GPU engine (collected last on purpose: the callback protocol itself is
covered on the CPU by tests/test_api_surface.py; this is the end-to-end flow that leaves ``best.ckpt`` for eval.py:169-172)."""
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
