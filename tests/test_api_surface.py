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


@pytest.mark.parametrize("kwargs", [{"context": "id", "input_dtype": "shuffled_set"}, {"context": "canvas", "input_dtype": "sorted_set"},
                                    {"seq_type": "flat"}, {"use_elemwise_noise": True}])
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


@pytest.mark.gpu
def test_shuffled_set_train_steps_track_the_oracle():
    """--input_dtype shuffled_set through the public train_step: shuffle -> corrupt -> PositionEmbedding encoder -> ... -> Adam,
    against the oracle replaying the same Philox streams."""
    from collections import OrderedDict

    from flex_dm_b200.mfp import MFP, Adam

    cols = make_input_columns("crello")
    m = MFP(cols, num_blocks=1, masking_method="random", input_dtype="shuffled_set", latent_dim=256, dropout=0.1, l2=1e-2, seed=17)
    m.set_weights(H.perturbed_weights(m.engine, 3))
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    assert "model/encoder/input_layer/const/embeddings/embeddings" in m.get_weights()
    o = O.OracleMFP(cols, num_blocks=1, masking_method="random", dropout=0.1, l2=1e-2, learning_rate=1e-3, clipnorm=1.0, input_dtype="shuffled_set")
    o.params = H.oracle_params_from_engine(m.engine)
    o.m = OrderedDict((k, torch.zeros_like(v)) for k, v in o.params.items())
    o.v = OrderedDict((k, torch.zeros_like(v)) for k, v in o.params.items())
    batch = make_synthetic_batch(cols, 4, 20, seed=6, lengths="ragged")
    for step in range(3):
        got = m.metrics_from_row(m.train_step(batch))
        ref = o.train_step(batch, seed=17, step=step)
        assert got["loss"] == pytest.approx(ref["loss"], rel=H.LOSS_RTOL), step


# ------------------------------------------------------------------------------------------------- eval.py::evaluate
def test_eval_mask_builders():
    """eval.py:52-93: elem mode repeats the document S times with eye(S) masks; group modes hide the group's fields on every
    valid element; the random mode raises as it does in the reference."""
    from flex_dm_b200.evaluation import build_masks

    cols = make_input_columns("crello")
    batch = make_synthetic_batch(cols, 1, 6, seed=3, fixed_lengths=np.asarray([6]))
    rep, masks = build_masks(batch, cols, "elem")
    assert rep["left"].shape == (6, 6, 1) and rep["length"].shape == (6, 1) and rep["image_embedding"].shape == (6, 6, 512)
    assert torch.equal(rep["left"][4], torch.as_tensor(batch["left"][0]))
    for key in ("type", "left", "color", "text_embedding"):
        assert torch.equal(masks[key], torch.eye(6, dtype=torch.bool)), key
    assert masks["canvas_width"].shape == (6,) and bool(masks["canvas_width"].all())
    ragged = make_synthetic_batch(cols, 3, 5, seed=4, fixed_lengths=np.asarray([5, 2, 4]))
    same, masks = build_masks(ragged, cols, "pos", ["left", "top", "width", "height"])
    seq = np.arange(5)[None, :] < np.asarray([5, 2, 4])[:, None]
    for key in ("left", "top", "width", "height"):
        assert np.array_equal(masks[key].numpy(), seq), key
    assert not bool(masks["type"].any()) and same["left"].shape == (3, 5, 1)
    with pytest.raises(TypeError):
        build_masks(batch, cols, "random")
    with pytest.raises(ValueError):
        build_masks(batch, cols, "attr")


@pytest.mark.gpu
@pytest.mark.parametrize("dataset,task_mode,group_keys", [("crello", "attr", ["opacity", "color", "font_family"]), ("rico", "pos", ["left", "top", "width", "height"]),
                                                          ("crello", "elem", None)])
def test_evaluate_matches_oracle_scores(dataset, task_mode, group_keys):
    """The per-field accuracies eval.py reports (score_num / score_den accumulated over batches), engine vs oracle."""
    from collections import OrderedDict, defaultdict

    from flex_dm_b200.evaluation import build_masks, evaluate
    from flex_dm_b200.mfp import MFP

    cols = make_input_columns(dataset)
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    m = MFP(cols, num_blocks=1, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=5)
    w = H.perturbed_weights(m.engine, 5)
    for name in w:  # sharper heads: accuracies that are neither 0 nor chance
        if name.startswith("model/decoder/") and name.endswith("/kernel"):
            w[name] = w[name] * 4.0
    m.set_weights(w)
    if task_mode == "elem":
        batches = [make_synthetic_batch(cols, 1, 7, seed=s, fixed_lengths=np.asarray([7])) for s in (1, 2)]
    else:
        batches = [make_synthetic_batch(cols, 4, 9, seed=s, lengths="ragged") for s in (1, 2)]
    got = evaluate(m, batches, cols, task_mode, (task_mode, group_keys) if group_keys else None, num_iter=1)
    params = H.oracle_params_from_engine(m.engine)
    total = defaultdict(float)
    for b in batches:
        example, masks = build_masks(b, cols, task_mode, group_keys)
        mod = O.preprocess_for_test(example, icols, masks)
        pred = O.merge_inputs_and_prediction(example, icols, masks, O.model_forward(params, mod, icols, 1))
        flag = torch.ones((example["left"].shape[0],), dtype=torch.bool) if (dataset == "rico" and task_mode == "pos") else None
        _, _, scores, _ = O.loss_layer(example, pred, masks, cols, flag)
        for k, v in scores.items():
            total[k] += float(v)
    assert set(got) == set(O.get_valid_input_columns(cols))
    for key, val in got.items():
        num, den = total[key + "_score_num"], total[key + "_score_den"]
        if den == 0.0:
            assert np.isnan(val), key
        elif icols[key]["type"] == "categorical":
            assert abs(val - num / den) <= max(1.0, 0.02 * den) / den, (key, val, num / den)  # argmax flips from TF32 near-ties
        else:
            assert val == pytest.approx(num / den, abs=5e-3), key
