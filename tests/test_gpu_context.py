"""GPU tests of ``--context id`` / ``--context length`` (reference: ``architecture/encoder.py:96-110,231-249``, ``decoder.py:74-78``,
``mfp.py:137``, ``eval.py:99-101``) through the public API: the host mirror pads the batch by one row for the context token, the engine
keeps the token in the row after each document's last element.  The reference-run goldens (``crello_ctx_id``, ``rico_ctx_length``) are
checked in ``test_golden_reference.py``; here full-length documents (the padding path), training curves, the demo call and the
evaluation loop are compared with the oracle."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from oracle import mfp_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _padded(batch):
    """The batch as the engine sees it: one all-padding row appended to every sequence column."""
    return {k: (np.pad(v, ((0, 0), (0, 1), (0, 0))) if v.ndim == 3 else v) for k, v in batch.items()}


@pytest.mark.parametrize("dataset,method,context", [("rico", "elem_pos_attr", "id"), ("crello", "random", "length"), ("crello", "elem_pos_attr_img_txt", "id"),
                                                    ("crello", "random", "canvas"), ("crello", "elem_pos_attr_img_txt", "canvas_add")])
def test_context_train_steps_track_the_oracle(dataset, method, context):
    from flex_dm_b200.mfp import MFP, Adam

    cols = make_input_columns(dataset)
    m = MFP(cols, num_blocks=2, masking_method=method, context=context, latent_dim=256, dropout=0.1, l2=1e-2, seed=5)
    m.set_weights(H.perturbed_weights(m.engine, seed=2))
    m.seed = 33
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    table = "model/encoder/input_layer/%s/embeddings" % {"id": "task", "length": "length"}.get(context, "format")
    rows = len(m.task_names) if context == "id" else 50 if context == "length" else cols["format"]["input_dim"] + 2
    assert table in m.get_weights() and m.get_weights()[table].shape == (rows, 256)
    assert ("model/decoder/decoders/format/kernel" in m.get_weights()) == (context == "canvas")  # decoder.py:25 use_canvas
    o = O.OracleMFP(cols, num_blocks=2, masking_method=method, dropout=0.1, l2=1e-2, dtype=torch.float64, learning_rate=1e-3, clipnorm=1.0, context=context)
    o.params = H.oracle_params_from_engine(m.engine)
    batch = make_synthetic_batch(cols, 5, 12, seed=1, lengths="ragged")  # holds a full-length document: exercises the extra row
    assert int(batch["length"].max()) == 11
    w0 = m.get_weights()
    for step in range(3):
        got = m.metrics_from_row(m.train_step(batch))
        ref = o.train_step(batch if context == "canvas_add" else _padded(batch), seed=33, step=step)
        assert got["loss"] == pytest.approx(ref["loss"], rel=H.LOSS_RTOL), step
        assert got["total_score"] == pytest.approx(ref["metrics"]["total_score"], abs=2e-2)
    w = m.get_weights()
    for name in (table, "model/blocks/seq2seq/seq2seq_0/attn/dense_value/kernel"):
        delta_ref = o.params[name].numpy() - w0[name]
        if name == table:  # rows of ids that never occurred only see the L2 term: compare the rows that got data gradients
            ids = np.unique(batch["length"][:, 0]) if context == "length" else np.arange(delta_ref.shape[0]) if context == "id" else np.unique(batch["format"][:, 0])
            assert H.rel_l2((w[name] - w0[name])[ids], delta_ref[ids]) < 0.15, name
        else:
            assert H.rel_l2(w[name] - w0[name], delta_ref) < 0.1, name


@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-simt", "tf32-tcgen05"])
def test_context_gradients_match_the_oracle(impl):
    """One backward pass with dropout off: the context table's gradient rows and an encoder table's <UNUSED> row (which the token's row
    must not leak into) against autograd."""
    from flex_dm_b200.mfp import MFP

    cols = make_input_columns("rico")
    m = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", context="id", latent_dim=256, dropout=0.0, l2=1e-2, seed=9)
    m.set_weights(H.perturbed_weights(m.engine, seed=4))
    eng = m.engine
    eng.set_gemm_impl(impl)
    batch = make_synthetic_batch(cols, 6, 8, seed=2, lengths="ragged")
    staged = m.stage(batch)
    B, S, length, dcols = m._bind(staged)
    assert S == 9
    tasks = torch.tensor([1, 3, 4, 3, 1, 4], dtype=torch.int32, device="cuda")
    m._set_context(tasks)
    eng.mask_corrupt(length, dcols, tasks, 7, 0)
    eng.forward(length, None, True, 7, 0)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True, sort_tasks=tasks)
    eng.backward(length, None, True, 7, 0)
    torch.cuda.synchronize()
    o = O.OracleMFP(cols, num_blocks=1, masking_method="elem_pos_attr", dropout=0.0, l2=None, dtype=torch.float64, context="id")
    o.params = H.oracle_params_from_engine(eng)
    inputs = o.to_torch(_padded(batch))
    targets, mod, masks = O.preprocess_for_train(inputs, o.input_columns, tasks.cpu(), O.PhiloxDraws(7, 0))
    r = o.step_from(targets, mod, masks, tasks.cpu(), None)
    assert float(row[3 * len(m.keys)].cpu()) == pytest.approx(r["data_loss"], rel=H.F32_LOSS_RTOL if impl == 1 else H.LOSS_RTOL)
    got = eng.get_weights(eng.grads)
    tol = H.F32_GRAD_REL_L2 if impl == 1 else H.GRAD_REL_L2
    for name in ("model/encoder/input_layer/task/embeddings", "model/encoder/input_layer/left/embeddings", "model/encoder/input_layer/type/embeddings",
                 "model/blocks/seq2seq/seq2seq_0/attn/dense_query/kernel", "model/decoder/decoders/icon/kernel"):
        assert H.rel_l2(got[name], r["grads"][name].numpy()) < tol, name
    unused_ids = [i for i in range(len(m.task_names)) if i not in (1, 3, 4)]
    assert np.all(got["model/encoder/input_layer/task/embeddings"][unused_ids] == 0.0)


def test_context_demo_call_and_evaluation():
    """``model(example, training=False, demo_args={"masks", "tasks"})`` and the eval.py loop with ``model.context == "id"``."""
    from flex_dm_b200.evaluation import evaluate
    from flex_dm_b200.mfp import MFP
    from flex_dm_b200.spec import get_attribute_groups

    cols = make_input_columns("rico")
    m = MFP(cols, num_blocks=1, masking_method="elem_pos_attr", context="id", latent_dim=256, dropout=0.1, l2=1e-2, seed=6)
    m.set_weights(H.perturbed_weights(m.engine, seed=6))
    B, S = 4, 10
    batch = make_synthetic_batch(cols, B, S, seed=8, lengths="ragged")
    t = H.to_torch(batch)
    seq = O.get_seq_mask(t["length"], S)
    masks = O.get_initial_masks(m.input_columns, seq)
    for key in ("left", "top", "width", "height"):
        masks[key] = seq
    task_id = m.task_names.index("pos")
    tasks = torch.full((B,), task_id, dtype=torch.int32)
    out = m(batch, training=False, demo_args={"masks": masks, "tasks": tasks})
    assert out["left"].shape == (B, S, 1, 64) and out["tasks"].tolist() == [task_id] * B
    # oracle: the same call on the padded batch, cropped
    p = H.oracle_params_from_engine(m.engine)
    pt = H.to_torch(_padded(batch))
    pmasks = OrderedDict((k, (torch.nn.functional.pad(v, (0, 1)) if v.dim() == 2 else v)) for k, v in masks.items())
    omod = O.preprocess_for_test(pt, m.input_columns, pmasks, tasks)
    ref = O.merge_inputs_and_prediction(pt, m.input_columns, pmasks, O.model_forward(p, omod, m.input_columns, 1, context="id"))
    for key in m.keys:
        assert np.abs(out[key].cpu().numpy() - ref[key][:, :S].numpy()).max() <= H.LOGIT_ATOL, key
    # another task id changes the prediction: the token is really read
    other = m(batch, training=False, demo_args={"masks": masks, "tasks": torch.full((B,), 1, dtype=torch.int32)})
    assert (other["left"] - out["left"]).abs().max() > 1e-4
    # inner boundary: Model.call on already-modified inputs reads modified_inputs["task"]
    inner = m.model({k: v.numpy() for k, v in omod.items() if k in m.input_columns or k == "task"}, training=False)
    oref = O.model_forward(p, omod, m.input_columns, 1, context="id")
    valid = seq.numpy()  # raw outputs at padded positions are never used; the row after a document's last element holds the token here
    for key in m.keys:
        assert np.abs(inner[key].cpu().numpy()[:, :S] - oref[key][:, :S].numpy())[valid].max() <= H.LOGIT_ATOL, key
    # eval.py loop
    group = ("pos", get_attribute_groups(cols.keys())["pos"])
    scores = evaluate(m, [batch], m.input_columns, "pos", group=group)
    assert all(0.0 <= scores[k] <= 1.0 for k in ("left", "top", "width", "height"))  # the other fields are 0/0 = nan, as in the reference
    # same numbers from the oracle's predictions pushed through the oracle's loss layer
    _, _, oscores, _ = O.loss_layer(pt, O.model_forward(p, omod, m.input_columns, 1, context="id"), pmasks, m.all_columns, torch.ones((B,), dtype=torch.bool))
    for k in ("left", "top", "width", "height"):
        want = float(oscores[k + "_score_num"]) / float(oscores[k + "_score_den"])
        assert scores[k] == pytest.approx(want, abs=0.08), k  # argmax ties under TF32 may flip single elements


def test_canvas_context_gradients_and_errors():
    """canvas / canvas_add: gradients of the canvas columns' tables against autograd (fp32 path), the never-read canvas heads of
    ``context="canvas"`` get exactly zero data gradient, and rico (no canvas columns) is refused like the reference's assertion."""
    from flex_dm_b200.mfp import MFP

    cols = make_input_columns("crello")
    for context in ("canvas", "canvas_add"):
        m = MFP(cols, num_blocks=1, masking_method="random", context=context, latent_dim=256, dropout=0.0, l2=1e-2, seed=9)
        m.set_weights(H.perturbed_weights(m.engine, seed=4))
        eng = m.engine
        eng.set_gemm_impl(1)
        batch = make_synthetic_batch(cols, 5, 7, seed=2, lengths="ragged")
        staged = m.stage(batch)
        B, S, length, dcols = m._bind(staged)
        assert S == (8 if context == "canvas" else 7)
        tasks = torch.zeros(B, dtype=torch.int32, device="cuda")
        m._set_context(tasks)
        eng.mask_corrupt(length, dcols, tasks, 7, 0)
        eng.forward(length, None, True, 7, 0)
        row = torch.zeros(eng.metrics_width, device="cuda")
        eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
        eng.backward(length, None, True, 7, 0)
        torch.cuda.synchronize()
        o = O.OracleMFP(cols, num_blocks=1, masking_method="random", dropout=0.0, l2=None, dtype=torch.float64, context=context)
        o.params = H.oracle_params_from_engine(eng)
        inputs = o.to_torch(_padded(batch) if context == "canvas" else batch)
        targets, mod, masks = O.preprocess_for_train(inputs, o.input_columns, tasks.cpu(), O.PhiloxDraws(7, 0))
        r = o.step_from(targets, mod, masks, tasks.cpu(), None)
        assert float(row[3 * len(m.keys)].cpu()) == pytest.approx(r["data_loss"], rel=H.F32_LOSS_RTOL)
        got = eng.get_weights(eng.grads)
        names = ["model/encoder/input_layer/%s/embeddings" % k for k in eng.canvas_keys]
        names += ["model/encoder/input_layer/left/embeddings", "model/encoder/input_layer/image_embedding/kernel", "model/blocks/seq2seq/seq2seq_0/attn/dense_key/kernel"]
        for name in names:
            assert H.rel_l2(got[name], r["grads"][name].numpy()) < H.F32_GRAD_REL_L2, (context, name)
        if context == "canvas":
            assert np.all(got["model/decoder/decoders/group/kernel"] == 0.0) and np.all(r["grads"]["model/decoder/decoders/group/kernel"].numpy() == 0.0)
    with pytest.raises(AssertionError):
        MFP(make_input_columns("rico"), num_blocks=1, context="canvas", latent_dim=256)
