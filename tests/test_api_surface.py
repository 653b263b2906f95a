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


@pytest.mark.parametrize("kwargs", [{"seq_type": "flat"}, {"use_elemwise_noise": True}])
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


# ================================================================================================= fit's callback protocol (CPU)
class _FitStub:
    """``MFP.fit`` with the engine taken out: the epoch runner is scripted, everything else (validation schedule, history, callback
    dispatch, stop_training) is the product code."""

    def __new__(cls, train_logs, val_logs):
        from flex_dm_b200.mfp import MFP

        class Stub(MFP):
            def __init__(self):  # no engine
                self.history, self.stop_training, self.saved = [], False, []
                self._train, self._val = iter(train_logs), iter(val_logs)

            def _run_epoch(self, iterator, steps, train, staged=False):
                from collections import OrderedDict
                return OrderedDict(next(self._train if train else self._val))

            def save_weights(self, path):
                self.saved.append(path)

        return Stub()


class _Dataset(list):
    yields_device_batches = True  # keeps fit from wrapping the (never read) dataset in a DevicePrefetcher


def test_fit_drives_the_reference_callback_list(tmp_path):
    """train.py:79-88 with ``callbacks=get_callbacks(args, dataspec, checkpoint_path)`` (helpers/callbacks.py:36-66): best.ckpt is
    written when ``val_total_score`` improves (mode max) and only on epochs that validated (validation_freq), per-epoch scalars are
    logged, hooks fire in Keras' order, plain callables keep working."""
    import json
    from types import SimpleNamespace

    from flex_dm_b200.callbacks import Callback, ModelCheckpoint, ScalarLogger, get_callbacks

    train = [{"loss": 5.0 - e, "total_score": 0.1 * e} for e in range(6)]
    val = [{"loss": 4.0, "total_score": 0.30}, {"loss": 3.0, "total_score": 0.50}, {"loss": 2.5, "total_score": 0.40}]
    model = _FitStub(train, val)
    events = []

    class Spy(Callback):
        def on_train_begin(self, logs=None):
            events.append("begin")

        def on_epoch_begin(self, epoch, logs=None):
            events.append("epoch_begin %d" % epoch)

        def on_epoch_end(self, epoch, logs=None):
            events.append("epoch_end %d %s" % (epoch, "val" if "val_total_score" in logs else "-"))

        def on_train_end(self, logs=None):
            events.append("end")

    seen = []
    args = SimpleNamespace(job_dir=str(tmp_path))
    os.makedirs(os.path.join(args.job_dir, "logs", "stale"))
    ckpt = os.path.join(args.job_dir, "checkpoints", "best.ckpt")
    cbs = get_callbacks(args, None, ckpt)
    assert [type(c).__name__ for c in cbs] == ["TensorBoard", "ModelCheckpoint", "TerminateOnNaN", "GarbageCollector"]
    scalars = ScalarLogger(os.path.join(args.job_dir, "logs"))  # the dependency-free logger next to it
    assert not os.path.exists(os.path.join(args.job_dir, "logs", "stale"))  # the log dir is overwritten, as in the reference
    history = model.fit(_Dataset(), steps_per_epoch=2, epochs=6, validation_data=_Dataset(), validation_steps=1, validation_freq=2,
                        callbacks=cbs + [scalars, Spy(), lambda epoch, logs, m: seen.append((epoch, m is model))], verbose=0)
    assert len(history) == 6 and [("val_loss" in h) for h in history] == [False, True] * 3
    assert model.saved == [ckpt, ckpt]  # epochs 2 (0.30) and 4 (0.50); epoch 6 (0.40) did not improve; odd epochs had nothing to monitor
    assert cbs[1].best == 0.50
    assert events == ["begin"] + [e for k in range(6) for e in ("epoch_begin %d" % k, "epoch_end %d %s" % (k, "val" if k % 2 else "-"))] + ["end"]
    assert seen == [(k, True) for k in range(6)]
    lines = [json.loads(x) for x in open(os.path.join(args.job_dir, "logs", "scalars.jsonl"))]
    assert [r["epoch"] for r in lines] == list(range(6)) and lines[1]["val_total_score"] == 0.30 and "val_loss" not in lines[0]
    # the TensorBoard callback wrote Keras' layout: epoch_<name> scalars, training under logs/train, val_* under logs/validation
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator

    for sub, want in (("train", [(e, 5.0 - e) for e in range(6)]), ("validation", [(1, 4.0), (3, 3.0), (5, 2.5)])):
        events = EventAccumulator(os.path.join(args.job_dir, "logs", sub))
        events.Reload()
        assert sorted(events.Tags()["scalars"]) == ["epoch_loss", "epoch_total_score"]
        assert [(e.step, e.value) for e in events.Scalars("epoch_loss")] == want
    with pytest.raises(NotImplementedError):
        ModelCheckpoint(ckpt)  # save_weights_only=False: a whole-model SavedModel has no counterpart
    assert ModelCheckpoint(ckpt, save_weights_only=True, monitor="val_acc").mode == "max" and ModelCheckpoint(ckpt, save_weights_only=True).mode == "min"


def test_fit_stops_on_request_and_on_a_non_finite_loss():
    from flex_dm_b200.callbacks import Callback, TerminateOnNaN

    class StopAfter(Callback):
        def on_epoch_end(self, epoch, logs=None):
            self.model.stop_training = epoch == 1

    ended = []

    class End(Callback):
        def on_train_end(self, logs=None):
            ended.append(True)

    model = _FitStub([{"loss": 1.0}] * 5, [])
    assert len(model.fit(_Dataset(), steps_per_epoch=1, epochs=5, callbacks=[StopAfter(), End()], verbose=0)) == 2 and ended == [True]
    model = _FitStub([{"loss": 1.0}, {"loss": float("nan")}, {"loss": 1.0}], [])
    history = model.fit(_Dataset(), steps_per_epoch=1, epochs=3, callbacks=[TerminateOnNaN(), End()], verbose=0)
    assert len(history) == 2 and np.isnan(history[-1]["loss"]) and ended == [True, True]
    stub = _FitStub([], [])
    cb = TerminateOnNaN()
    cb.set_model(stub)
    cb.on_epoch_end(0, {"loss": float("inf")})
    assert stub.stop_training


def test_notebook_imports_resolve_and_merge_matches_the_oracle():
    """The names the reference notebooks import (demo_rico / demo_crello cell 1): ``merge_inputs_and_prediction`` from the MFP module,
    ``get_seq_mask`` / ``get_initial_masks`` from masking, ``ATTRIBUTE_GROUPS`` / ``DataSpec`` / ``set_visual_default`` from the data
    module.  The host-level merge equals the oracle's restatement of mfp.py:46-69, is idempotent (the notebook applies it to an already
    merged prediction) and copies demo-only columns."""
    from flex_dm_b200.dataspec import ATTRIBUTE_GROUPS, DataSpec, set_visual_default  # noqa: F401
    from flex_dm_b200.masking import get_initial_masks, get_seq_mask
    from flex_dm_b200.mfp import MFP, merge_inputs_and_prediction  # noqa: F401

    assert sorted(ATTRIBUTE_GROUPS["crello"]) == ["attr", "img", "pos", "txt", "type"]
    doc = set_visual_default({"elements": [{"color": [1.0, 0.5, 0.0], "opacity": 0.3, "font_family": "X", "left": 0.1}]})
    assert doc["elements"][0] == {"color": [0.0, 0.0, 0.0], "opacity": 1.0, "font_family": "DummyFont", "left": 0.1}
    cols = make_input_columns("crello")
    assert cols["id"].get("demo_only") and cols["uuid"].get("demo_only")  # strings, demo only (crello-spec.yml)
    batch = make_synthetic_batch(cols, 3, 6, seed=4, lengths="ragged")
    inputs = {k: torch.as_tensor(v) for k, v in batch.items()}
    inputs["id"] = np.array([[b"doc%d" % b] for b in range(3)], dtype=object)
    inputs["uuid"] = np.array([[b"a"] * 6] * 3, dtype=object)
    seq_mask = get_seq_mask(inputs["length"])
    masks = get_initial_masks({k: v for k, v in cols.items() if not v.get("demo_only")}, seq_mask)
    gen = torch.Generator().manual_seed(0)
    for key in ("left", "color", "image_embedding"):
        masks[key] = (torch.rand(seq_mask.shape, generator=gen) < 0.5) & seq_mask
    prediction = {}
    for key, c in cols.items():
        if c["is_sequence"] and not c.get("demo_only"):
            shape = tuple(inputs[key].shape) + ((c["input_dim"],) if c["type"] == "categorical" else ())
            prediction[key] = torch.randn(shape, generator=gen)
    ref = O.merge_inputs_and_prediction(inputs, cols, masks, {k: v.clone() for k, v in prediction.items()})
    got = merge_inputs_and_prediction(inputs, cols, masks, prediction)
    assert got is prediction and got["uuid"] is inputs["uuid"] and got["id"] is inputs["id"]
    assert set(got) == set(ref)
    for key in ref:
        if key not in ("id", "uuid"):
            assert torch.equal(torch.as_tensor(got[key]), torch.as_tensor(ref[key])), key
    assert torch.equal(got["left"].argmax(-1)[~masks["left"]], inputs["left"][~masks["left"]])  # one-hot ground truth where not masked
    again = merge_inputs_and_prediction(inputs, cols, masks, {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in got.items()})
    for key in ref:
        if key not in ("id", "uuid"):
            assert torch.equal(torch.as_tensor(again[key]), torch.as_tensor(got[key])), key


def _train_args(tmp_path, name, **over):
    from types import SimpleNamespace

    args = dict(dataset_name=name, data_dir=str(tmp_path / "data"), job_dir=str(tmp_path / "job"), batch_size=8, weights=None, latent_dim=256,
                num_blocks=1, arch_type="oneshot", block_type="deepsvg", l2=1e-2, dropout=0.1, masking_method="random", seq_type="default",
                context=None, input_dtype="set", learning_rate=1e-3, num_epochs=3, validation_freq=5, verbose=0, seed=4)
    args.update(over)
    return SimpleNamespace(**args)


def test_train_function_drives_the_job_like_the_reference(tmp_path, monkeypatch):
    """``training.train(args)`` (train.py:16-97) with the model replaced by a recorder: the dataset / model / fit / evaluate / save calls
    and their arguments are the reference's, and the job directory gets args.json, logs/ and checkpoints/final.ckpt."""
    import json

    from flex_dm_b200 import training
    from flex_dm_b200.synthetic import write_synthetic_dataset

    write_synthetic_dataset(str(tmp_path / "data"), "rico", {"train": 24, "val": 8, "test": 8}, seq_len=6, shards=1, seed=5)
    calls = []

    class Recorder:
        metrics_names = ["loss", "total_score"]

        def __init__(self, input_columns, **kwargs):
            calls.append(("init", sorted(kwargs.items()), "left" in input_columns))

        def load_weights(self, path):
            calls.append(("load", path))

        def compile(self, optimizer=None, **kw):
            calls.append(("compile", optimizer.learning_rate, optimizer.clipnorm))

        def fit(self, dataset, **kw):
            first = next(iter(dataset))
            calls.append(("fit", {k: v for k, v in kw.items() if k not in ("validation_data", "callbacks")}, tuple(first["left"].shape[:1]),
                          [type(c).__name__ for c in kw["callbacks"]], len(list(kw["validation_data"]))))

        def evaluate(self, dataset, batch_size=None):
            calls.append(("evaluate", len(list(dataset)), batch_size))
            return [1.5, 0.25]

        def save_weights(self, path):
            calls.append(("save", path))

    monkeypatch.setattr(training, "MFP", Recorder)
    args = _train_args(tmp_path, "rico", weights="/some/init.ckpt")
    results = training.train(args)
    assert results == {"loss": 1.5, "total_score": 0.25}
    job = args.job_dir
    assert json.load(open(os.path.join(job, "args.json")))["masking_method"] == "random"
    want_kwargs = sorted(dict(num_blocks=1, block_type="deepsvg", masking_method="random", seq_type="default", arch_type="oneshot", context=None,
                              latent_dim=256, dropout=0.1, l2=1e-2, input_dtype="set", seed=4).items())
    assert calls == [
        ("init", want_kwargs, True),
        ("load", "/some/init.ckpt"),
        ("compile", 1e-3, 1.0),
        ("fit", dict(steps_per_epoch=3, epochs=3, validation_steps=1, validation_freq=3, verbose=0), (8,),
         ["TensorBoard", "ModelCheckpoint", "TerminateOnNaN", "GarbageCollector"], 1),
        ("evaluate", 1, 8),
        ("save", os.path.join(job, "checkpoints", "final.ckpt")),
    ]
