"""The Keras-facing surface the reference's callers use: constructor keywords and their errors, ``compile`` /
``fit`` / ``evaluate`` / ``metrics_names`` (train.py:53-97), ``save_weights`` / ``load_weights`` (train.py:67-69,94-97), and the
long-sequence path of the engine."""
import itertools
import os

import numpy as np
import pytest
import torch

from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from oracle import mfp_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("kwargs", [{"block_type": "transformer"}, {"context": "id"}, {"input_dtype": "shuffled_set"}, {"seq_type": "flat"},
                                    {"use_elemwise_noise": True}])
def test_unsupported_switches_raise_instead_of_being_ignored(kwargs):
    from flex_dm_b200.mfp import MFP

    with pytest.raises(NotImplementedError):
        MFP(make_input_columns("crello"), **kwargs)


@pytest.mark.gpu
def test_fit_evaluate_and_checkpoint_roundtrip(tmp_path):
    """train.py:53-97 end to end on a tiny model: a few epochs reduce the loss, evaluate() follows metrics_names, and a saved
    checkpoint restores the same weights and the same evaluation."""
    from flex_dm_b200.mfp import MFP, Adam

    cols = make_input_columns("rico")
    model = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", latent_dim=256, dropout=0.1, l2=1e-2, seed=1)
    model.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0), run_eagerly=True)
    batches = [make_synthetic_batch(cols, 8, 12, seed=i, lengths="ragged") for i in range(3)]
    history = model.fit(itertools.cycle(batches), steps_per_epoch=6, epochs=3, validation_data=batches[:1], validation_steps=1, verbose=0)
    assert len(history) == 3 and history[-1]["loss"] < history[0]["loss"]
    assert {"loss", "total_score", "val_loss", "left_loss", "type_score"} <= set(history[-1])
    names = model.metrics_names
    assert names[0] == "loss" and "total_score" in names
    model._step = 1000  # evaluation corrupts inputs with the step-indexed Philox streams: fix the step to compare two runs
    values = model.evaluate(batches)
    assert len(values) == len(names) and all(np.isfinite(values))
    path = os.path.join(tmp_path, "checkpoints", "final.ckpt")
    model.save_weights(path)
    other = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", latent_dim=256, dropout=0.1, l2=1e-2, seed=99)
    other.compile(optimizer="adam")
    other.load_weights(path)
    for k, v in model.get_weights().items():
        assert np.array_equal(v, other.get_weights()[k]), k
    other.seed, other._step = model.seed, 1000  # same Philox key and step -> same corruption
    assert other.evaluate(batches) == pytest.approx(values, rel=1e-6)


@pytest.mark.gpu
def test_sequences_longer_than_the_tensor_core_tile():
    """S > 128 elements: the product path keeps its tcgen05 GEMMs and routes the attention core through the SIMT kernels."""
    from flex_dm_b200.mfp import MFP

    cols = make_input_columns("rico", max_length=160)
    m = MFP(cols, num_blocks=1, masking_method="random", latent_dim=256, dropout=0.0, l2=1e-2, seed=2)
    m.set_weights(H.perturbed_weights(m.engine, 2))
    B, S = 2, 160
    batch = make_synthetic_batch(cols, B, S, seed=4, fixed_lengths=np.asarray([160, 131]))
    got = m.model(batch, training=False)
    params = H.oracle_params_from_engine(m.engine)
    ref = O.model_forward(params, {k: torch.as_tensor(v) for k, v in batch.items()}, m.input_columns, 1)
    for key in m.keys:
        assert np.abs(got[key].cpu().numpy() - ref[key].numpy()).max() <= H.LOGIT_ATOL, key
