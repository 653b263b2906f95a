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


def test_tf32_operand_rounding_of_the_product_path_is_recorded():
    """What the tcgen05 GEMM does to fp32 operands below TF32's 10-bit mantissa.  DESIGN.md section 2 states round-to-nearest (the TMA
    unit converts TFLOAT32 maps; the MMA itself would truncate); the oracle's TF32 emulation (oracle.emulate_tf32) and the ReLU-gate
    analysis of tests/test_golden_reference.py depend on which it is.  The products below are exact in fp32 under either rule, so the
    result identifies the rule; it is reported as a warning and only garbage fails.  (Written without GPU time: not yet run on a B200.)"""
    import warnings

    import torch

    from flex_dm_b200.engine import debug_gemm

    up, half, lo = 2.0 ** -10, 2.0 ** -11, 2.0 ** -13
    cases = {"above half an ulp": 1.0 + half + lo, "tie": 1.0 + half, "below half an ulp": 1.0 + half - lo, "negative, above half": -(1.0 + half + lo)}
    seen = {}
    for operand in ("A", "B"):  # the value sits in the A (token-major activations) or in the B (weights) operand
        for label, value in cases.items():
            M = N = 128
            K = 32
            A = torch.zeros(M, K)
            Bm = torch.zeros(N, K)
            A[:, 0] = value if operand == "A" else 1.0
            Bm[:, 0] = 1.0 if operand == "A" else value
            out = debug_gemm(A.cuda(), 0, Bm.T.contiguous().cuda(), 1, M, N, K, impl=0)  # forward layout: x . W, W stored [K][N]
            torch.cuda.synchronize()
            got = out.cpu()
            assert bool((got == got[0, 0]).all())
            seen[(operand, label)] = float(got[0, 0])
    rules = set()
    for (operand, label), got in seen.items():
        sign = -1.0 if label.startswith("negative") else 1.0
        assert got in (sign * 1.0, sign * (1.0 + up)), (operand, label, got)
    for operand in ("A", "B"):
        above, tie, below = (seen[(operand, k)] for k in ("above half an ulp", "tie", "below half an ulp"))
        rule = "truncation" if above == 1.0 else ("nearest, ties away" if tie == 1.0 + up else "nearest, ties to even")
        assert below == 1.0 and (seen[(operand, "negative, above half")] == -above)
        rules.add((operand, rule))
    warnings.warn("TF32 operand rounding of the tcgen05 GEMM path: %s" % ", ".join("%s operand: %s" % r for r in sorted(rules)))


@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-simt", "tf32-tcgen05"])
@pytest.mark.xfail(reason="engine path written without GPU time; not yet run on a B200", strict=False)
def test_engine_cases_not_yet_run_on_a_gpu(impl):
    """The reference-run goldens whose engine path was written after the round's GPU time had run out (tests/test_golden_reference.py::
    UNVERIFIED: --context id with --input_dtype shuffled_set), through the same checks as every other golden case -- collected last so that
    never-run kernel code cannot disturb the verified tests; an XPASS here is the confirmation that lets MFP.allow_unverified go."""
    from tests.test_golden_reference import UNVERIFIED
    from tests.test_golden_reference import test_engine_matches_reference_python as check

    for case in sorted(UNVERIFIED):
        check(case, impl)
