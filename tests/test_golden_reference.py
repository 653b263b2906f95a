"""Golden vectors produced by the REFERENCE'S OWN PYTHON (tests/golden/make_golden.py runs /root/reference's
MFP.call / preprocess_for_train / Model.call / LossLayer.call on the torch-backed TensorFlow stand-in,
oracle/tf_standin/) against (a) the CPU oracle -- CPU tests, this is what pins the oracle -- and (b) the CUDA
engine through the C ABI -- GPU tests.  The .npz fixtures are committed; nothing here reads /root/reference.

Masking outputs (task selection, <MASK>/<UNUSED>/random tokens, per-field masks) are compared bit-exactly; floating
point within the tolerances of tests/helpers.py."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from flex_dm_b200.spec import make_input_columns
from oracle import mfp_oracle as O
from tests import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WEIGHT_SEED, RATE, L2, LR = 11, 0.1, 1e-2, 1e-4  # as in make_golden.py
CASES = {  # name: (dataset, masking_method, num_blocks, seed, step)
    "crello_random": ("crello", "random", 2, 7, 0),
    "crello_multi": ("crello", "elem_pos_attr_img_txt", 2, 3, 2),
    "rico_pos": ("rico", "elem_pos_attr", 2, 5, 1),
    "crello_postln": ("crello", "random", 2, 9, 3),  # --block_type transformer (post-LayerNorm block)
    "rico_shuffled": ("rico", "random_elem_pos_attr", 2, 13, 5),  # --input_dtype shuffled_set (shuffle + PositionEmbedding)
    "crello_sorted": ("crello", "random", 2, 15, 1),  # --input_dtype sorted_set (sort_inputs + PositionEmbedding)
    # --context id / length (encoder.py:96-110,231-249): the batch keeps one free row per document for the context token; the
    # reference ran on the same batch cut to its longest document, so its arrays have one column less (``cut``)
    "crello_ctx_id": ("crello", "elem_pos_attr_img_txt", 2, 21, 2),
    "rico_ctx_length": ("rico", "elem_pos_attr", 2, 30, 1),
    # --context canvas (token = sum of the canvas columns' embeddings) / canvas_add (that sum added to every element; no token)
    "crello_ctx_canvas": ("crello", "random", 2, 49, 0),
    "crello_ctx_canvas_add": ("crello", "elem_pos_attr_img_txt", 2, 27, 1),
    # --context id with --input_dtype shuffled_set: positions are added after the token was put in front (encoder.py:247-252; engine:
    # PosEmbed::shift, pos_embed_bwd_ctx_kernel)
    "rico_ctx_id_shuffled": ("rico", "random_elem_pos_attr", 2, 37, 2),
    "crello_ctx_length_sorted": ("crello", "random", 2, 33, 1),  # ... and --context length with --input_dtype sorted_set
}
BLOCK_TYPE = {"crello_postln": "transformer"}
INPUT_DTYPE = {"rico_shuffled": "shuffled_set", "crello_sorted": "sorted_set", "rico_ctx_id_shuffled": "shuffled_set",
               "crello_ctx_length_sorted": "sorted_set"}
CONTEXT = {"crello_ctx_id": "id", "rico_ctx_length": "length", "crello_ctx_canvas": "canvas", "crello_ctx_canvas_add": "canvas_add",
           "rico_ctx_id_shuffled": "id", "crello_ctx_length_sorted": "length"}
TOKEN_CASES = {c for c, ctx in CONTEXT.items() if ctx != "canvas_add"}  # cases whose batch keeps a free row for the context token


def cut(x, case):
    """Sequence arrays of a context case without the engine batch's free last row (what the reference saw)."""
    return x[:, :-1] if case in TOKEN_CASES else x


def projection_vector(name, n):  # same as make_golden.py
    h = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") ^ (len(name) * 0x9E3779B97F4A7C15 & (2**63 - 1))
    rng = np.random.Generator(np.random.PCG64(h))
    return rng.integers(0, 2, size=n).astype(np.float64) * 2.0 - 1.0


def load(case):
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    dataset, method, L, seed, step = CASES[case]
    batch = OrderedDict((k[3:], g[k]) for k in g.files if k.startswith("in/"))
    cols = make_input_columns(dataset, max_length=50)
    return g, cols, batch, method, L, seed, step


def test_golden_files_cover_every_task_and_edge_case():
    """The fixtures exercise what the reference's masking/loss code branches on: every task id of both datasets,
    a single-element document, full-length documents, the rico sort branch."""
    seen = set()
    for case in CASES:
        g = np.load(os.path.join(GOLDEN, case + ".npz"))
        seen |= {(CASES[case][0], int(t)) for t in g["tasks"]}
        assert int(g["in/length"].min()) == 0 and int(g["in/length"].max()) == g["in/left"].shape[1] - (2 if case in TOKEN_CASES else 1)
    assert {t for d, t in seen if d == "crello"} == {0, 1, 3, 4, 5, 6}
    assert "crello_postln" in CASES  # the post-LayerNorm block of --block_type transformer
    assert {t for d, t in seen if d == "rico"} >= {1, 3, 4}


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_matches_reference_python(case):
    g, cols, batch, method, L, seed, step = load(case)
    input_dtype = INPUT_DTYPE.get(case, "set")
    o = O.OracleMFP(cols, num_blocks=L, masking_method=method, dropout=RATE, l2=L2, learning_rate=LR, clipnorm=1.0,
                    block_type=BLOCK_TYPE.get(case, "deepsvg"), input_dtype=input_dtype, context=CONTEXT.get(case))
    o.params = O.init_params(cols, L, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05, input_dtype=input_dtype, context=CONTEXT.get(case))
    draws = O.PhiloxDraws(seed, step)
    tasks = torch.as_tensor(g["tasks"])
    assert set(g["tasks"].tolist()) <= set(o.allowed_tasks)
    inputs = o.to_torch(batch)
    targets, mod, masks = O.preprocess_for_train(inputs, o.input_columns, tasks, draws, input_dtype)
    if input_dtype != "set":  # the batch the reference's shuffle_inputs / sort_inputs produced (tensor_utils.py:14-76)
        for key, column in o.input_columns.items():
            assert np.array_equal(cut(targets[key].numpy(), case) if column["is_sequence"] else targets[key].numpy(), g["tgt/" + key]), key
    # ---- masking path: bit-exact against the reference's preprocess_for_train (mfp.py:95-138)
    for key, column in o.input_columns.items():
        seq = column["is_sequence"]
        assert np.array_equal(cut(mod[key].numpy(), case) if seq else mod[key].numpy(), g["mod/" + key]), key
        assert np.array_equal(cut(masks[key].numpy(), case) if seq else masks[key].numpy(), g["mask/" + key]), key
        if seq and case in TOKEN_CASES:
            assert not masks[key][:, -1].any(), key  # the free row is padding
    assert np.array_equal(mod["task"].numpy(), g["mod/task"])
    B, S = batch["left"].shape[:2]
    r = o.step_from(targets, mod, masks, tasks, o.dropout_masks(draws, B, S))
    # ---- Model.call (model.py:26-30): raw logits
    for key, v in r["outputs"].items():
        assert np.abs(cut(v.detach().numpy(), case) - g["logits/" + key]).max() <= 1e-10, key
    # ---- LossLayer.call (metrics.py:173-299) + regularisers
    assert r["loss"] == pytest.approx(float(g["total_loss"]), rel=1e-12)
    assert r["data_loss"] == pytest.approx(float(g["data_loss"]), rel=1e-12)
    for key, v in r["losses"].items():
        assert v == pytest.approx(float(g["loss/" + key]), rel=1e-11, abs=1e-12), key
    for key, v in r["scores"].items():
        k, what = key.rsplit("_score_", 1)
        assert v == pytest.approx(float(g["score_%s/%s" % (what, k)]), rel=1e-11, abs=1e-12), key
    for key, v in r["metrics"].items():
        if not key.endswith("_loss"):
            assert v == pytest.approx(float(g["metric/" + key]), rel=1e-11, abs=1e-12), key
    # ---- gradients of the total loss and the variables after Adam(1e-4, clipnorm=1) (train.py:71-77)
    for name, gr in r["grads"].items():
        gr = gr.numpy().reshape(-1)
        scale = max(float(g["gradnorm/" + name]), 1e-12)
        assert abs(np.linalg.norm(gr) - float(g["gradnorm/" + name])) <= 1e-9 * scale, name
        assert abs(gr @ projection_vector(name, gr.size) - float(g["gradproj/" + name])) <= 1e-9 * scale * np.sqrt(gr.size), name
        assert np.abs(gr[:8] - g["gradhead/" + name]).max() <= 1e-9 * scale, name
        w = o.params[name].numpy().reshape(-1)
        assert np.abs(w[:8] - g["newhead/" + name]).max() <= 1e-12, name
        assert abs(w @ projection_vector(name, w.size) - float(g["newproj/" + name])) <= 1e-10 * np.sqrt(w.size), name


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_merge_matches_reference_python(case):
    """MFP.call's return value (mfp.py:46-69,342-347) from the float32 run of the reference."""
    g, cols, batch, method, L, seed, step = load(case)
    input_dtype = INPUT_DTYPE.get(case, "set")
    o = O.OracleMFP(cols, num_blocks=L, masking_method=method, dropout=RATE, l2=L2, dtype=torch.float32, input_dtype=input_dtype, context=CONTEXT.get(case))
    o.params = O.init_params(cols, L, 256, WEIGHT_SEED, torch.float32, bias_scale=0.05, input_dtype=input_dtype, context=CONTEXT.get(case))
    draws = O.PhiloxDraws(seed, step)
    tasks = torch.as_tensor(g["tasks"])
    inputs = o.to_torch(batch)
    targets, mod, masks = O.preprocess_for_train(inputs, o.input_columns, tasks, draws, input_dtype)
    B, S = batch["left"].shape[:2]
    outputs = O.model_forward(o.params, mod, o.input_columns, L, o.dropout_masks(draws, B, S), RATE, block_type=BLOCK_TYPE.get(case, "deepsvg"),
                              context=CONTEXT.get(case))
    merged = O.merge_inputs_and_prediction(inputs, o.input_columns, masks, outputs)
    keys = [k[7:] for k in g.files if k.startswith("merged/")]
    assert set(keys) == set(o.input_columns)
    for key in keys:
        ref = g["merged/" + key]
        got = merged[key].detach().numpy()
        if o.input_columns[key]["is_sequence"]:
            got = cut(got, case)
        assert got.shape == ref.shape, key
        if o.input_columns[key]["is_sequence"]:
            assert np.abs(got - ref).max() <= 2e-4, key
        else:
            assert np.array_equal(got, ref), key


# ================================================================================================= oracle runs with hooks
# (tools/tf32_gate_scan.py uses these to check that no fixture hinges on a single ReLU gate inside TF32's rounding error: round 1's
# --context canvas fixture did -- one gate carried 10.8 % of a gradient's norm and either side is a correct TF32 result -- and was
# regenerated with another seed instead of teaching the test to accept two answers.)
def oracle_step(case, closed_gates=(), tf32=None, record=None):
    """The oracle's train step on a fixture -> (grads dict of numpy, params after Adam); ``closed_gates`` forces single ReLU gates of the
    FFNs shut, ``tf32`` = an operand-rounding function to run every matrix product in emulated TF32, ``record`` = dict that receives
    the FFN pre-activations per block."""
    g, cols, batch, method, L, seed, step = load(case)
    input_dtype = INPUT_DTYPE.get(case, "set")
    o = O.OracleMFP(cols, num_blocks=L, masking_method=method, dropout=RATE, l2=L2, learning_rate=LR, clipnorm=1.0,
                    block_type=BLOCK_TYPE.get(case, "deepsvg"), input_dtype=input_dtype, context=CONTEXT.get(case))
    o.params = O.init_params(cols, L, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05, input_dtype=input_dtype, context=CONTEXT.get(case))
    draws = O.PhiloxDraws(seed, step)
    tasks = torch.as_tensor(g["tasks"])
    targets, mod, masks = O.preprocess_for_train(o.to_torch(batch), o.input_columns, tasks, draws, input_dtype)
    B, S = batch["left"].shape[:2]

    def hook(i, pre):
        if record is not None:
            record[i] = pre.detach().clone()
        keep = torch.ones_like(pre)
        for block, index in closed_gates:
            if block == i:
                keep[index] = 0.0
        return torch.relu(pre) * keep

    saved_hook = O._relu_hook
    O._relu_hook = hook
    try:
        if tf32 is None:
            r = o.step_from(targets, mod, masks, tasks, o.dropout_masks(draws, B, S))
        else:
            with O.emulate_tf32(tf32):
                r = o.step_from(targets, mod, masks, tasks, o.dropout_masks(draws, B, S))
    finally:
        O._relu_hook = saved_hook
    return OrderedDict((k, v.numpy()) for k, v in r["grads"].items()), OrderedDict((k, v.detach().numpy()) for k, v in o.params.items())


def summaries(grads, params):
    """The per-variable summaries a golden file keeps (make_golden.py): norm, +-1 projection and first 8 entries of the gradient, first 8
    entries of the updated variable."""
    out = {}
    for name, gr in grads.items():
        gr = gr.reshape(-1)
        out["gradnorm/" + name] = np.linalg.norm(gr)
        out["gradproj/" + name] = gr @ projection_vector(name, gr.size)
        out["gradhead/" + name] = gr[:8].copy()
        out["newhead/" + name] = params[name].reshape(-1)[:8].copy()
    return out


tf32_truncate = O.tf32_truncate


def check_gradient_summaries(total_grads, g, grad_tol):
    """Gradients of the total loss (data + L2), flat float64 per variable, against the summaries of a reference run."""
    for name, gg in total_grads.items():
        scale = max(float(g["gradnorm/" + name]), 1e-9)
        if name.endswith("dense_key/bias"):
            # softmax is shift-invariant: the exact data gradient of the key bias is 0 (what is left in the golden value is
            # the L2 term); the engine's is the rounding noise of the column sum of dK -- bound it by the kernel's gradient
            scale = max(scale, float(g["gradnorm/" + name.replace("/bias", "/kernel")]))
        assert abs(np.linalg.norm(gg) - float(g["gradnorm/" + name])) <= grad_tol * scale, name
        assert abs(gg @ projection_vector(name, gg.size) - float(g["gradproj/" + name])) <= grad_tol * scale * 4, name
        assert np.abs(gg[:8] - g["gradhead/" + name]).max() <= grad_tol * scale, name


def check_adam_heads(new_weights, g, impl):
    """The variables after L2 + per-variable clipnorm + one Adam step against a reference run."""
    for name, w in new_weights.items():
        # The first Adam step is lr * g / (|g| + 3.2e-6) on the clipped gradient: entries far above that scale move by
        # exactly +-lr (compared strictly); entries whose exact gradient is ~0 move by up to lr in the direction of the
        # rounding noise, in the engine and in a float32 TensorFlow run alike (bounded by 2 lr).
        gref = g["gradhead/" + name] * min(1.0, 1.0 / max(float(g["gradnorm/" + name]), 1e-12))
        # (TF32: a clipped entry of 1e-3 is 1e-3 of the gradient's norm, the size of the product path's own rounding error)
        strict = np.abs(gref) > (1e-3 if impl != 0 else 5e-3)
        diff = np.abs(w[:8] - g["newhead/" + name])
        assert diff[strict].max(initial=0.0) <= 5e-6, name
        assert diff.max() <= 2.0 * LR + 5e-6, name


def _flat(d):
    return OrderedDict((k, np.asarray(v, dtype=np.float64).reshape(-1)) for k, v in d.items())


@pytest.mark.parametrize("case", list(CASES))
def test_engine_checks_accept_the_oracle_run(case):
    """The comparison helpers of the GPU test below, exercised on the CPU: the oracle's own gradients and Adam step pass them at the
    fp32-path tolerances against every golden file (so a failure on the GPU is the engine's, not the helpers')."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    grads, params = oracle_step(case)
    check_gradient_summaries(_flat(grads), g, H.F32_GRAD_REL_L2)
    check_adam_heads(_flat(params), g, 1)


@pytest.mark.parametrize("case", list(CASES))
def test_fixtures_are_stable_under_tf32_rounding(case):
    """Every fixture must be a fair target for the TF32 product path: the oracle's TF32 emulation (round-to-nearest-even operands, the rule
    measured on the tcgen05 path by tests/test_gpu_parity.py::test_tf32_operand_rounding_of_the_product_path) passes the GPU test's checks at
    the product path's tolerances.  A fixture that fails here hinges on a discrete event inside TF32's rounding error -- a ReLU gate at zero
    (tools/tf32_gate_scan.py), or two predicted elements of a rico ``pos`` document swapping places in the argmax sort of the loss
    (tensor_utils.py:14-44) -- and gets another seed in tests/golden/make_golden.py instead of a test that accepts two answers."""
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    rna, rna_params = oracle_step(case, tf32=O.tf32_round)
    check_gradient_summaries(_flat(rna), g, H.GRAD_REL_L2)
    check_adam_heads(_flat(rna_params), g, 0)


# ================================================================================================= GPU (C ABI)
def _engine_cases():
    for case in CASES:
        for impl, name in ((1, "fp32-simt"), (0, "tf32-tcgen05"), (2, "3xtf32-tcgen05")):
            yield pytest.param(case, impl, id="%s-%s" % (case, name))


@pytest.mark.gpu
@pytest.mark.parametrize("case,impl", list(_engine_cases()))
def test_engine_matches_reference_python(case, impl):
    from flex_dm_b200.mfp import MFP

    g, cols, batch, method, L, seed, step = load(case)
    input_dtype = INPUT_DTYPE.get(case, "set")
    m = MFP(cols, num_blocks=L, block_type=BLOCK_TYPE.get(case, "deepsvg"), masking_method=method, input_dtype=input_dtype, latent_dim=256,
            dropout=RATE, l2=L2, seed=0, context=CONTEXT.get(case))
    m._pad_context = False  # the golden batches already keep one free row per document (see CASES)
    params = O.init_params(cols, L, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05, input_dtype=input_dtype, context=CONTEXT.get(case))
    m.set_weights({k: v.numpy().astype(np.float32) for k, v in params.items()})
    eng = m.engine
    eng.set_gemm_impl(impl)
    # impl 2 (compensated 3xTF32 on the tensor cores) is held to the fp32 tolerances, like the SIMT fp32 path
    logit_atol, loss_rtol, grad_tol = H.tolerances(impl)
    B, S = batch["left"].shape[:2]
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    tasks = torch.as_tensor(g["tasks"]).cuda()
    m._set_context(tasks)
    if input_dtype != "set":  # shuffle_inputs / sort_inputs: permutation and reordered columns bit-exact against the reference's
        perm = torch.zeros((B, S), dtype=torch.int32, device="cuda")
        dcols = eng.shuffle_inputs(length, dcols, seed, step, perm_out=perm)
        torch.cuda.synchronize()
        assert np.array_equal(perm.cpu().numpy(), g["perm"])
        for f, key in enumerate(m.keys):
            assert np.array_equal(cut(dcols[f].cpu().numpy(), case), g["tgt/" + key]), key
    eng.mask_corrupt(length, dcols, tasks, seed, step)
    torch.cuda.synchronize()
    # ---- masking: bit-exact against the reference's preprocess_for_train
    for f, key in enumerate(m.keys):
        assert np.array_equal(cut(eng.masks[f].cpu().numpy().astype(bool), case), g["mask/" + key]), key
        got = cut(eng.modified[f].cpu().numpy(), case)
        if cols[key]["type"] == "categorical":
            assert np.array_equal(got, g["mod/" + key]), key
        else:
            # <MASK> (10.0) / <UNUSED> (0.0) rows exactly; random replacements are Box-Muller normals whose device
            # logf/cosf differ from numpy's in the last bits
            ref = g["mod/" + key].astype(np.float32)
            assert np.array_equal(got == 10.0, ref == 10.0) and np.array_equal(got == 0.0, ref == 0.0), key
            assert np.allclose(got, ref, atol=1e-6, rtol=0), key
    # ---- forward (training, Philox dropout) / loss / backward
    logits = torch.empty((B * S, eng.logit_width), device="cuda")
    eng.forward(length, None, True, seed, step, logits_out=logits)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True, sort_tasks=tasks if m.sort_pos else None)
    eng.backward(length, None, True, seed, step)
    torch.cuda.synchronize()
    got = m.split_logits(logits, B, S)
    valid = np.ones((B, S), dtype=bool)
    if case in TOKEN_CASES:
        # the engine keeps the context token in the first padding row of each document, where the reference computes a (never used)
        # prediction for a padded element: raw logits are comparable on the documents' own elements
        valid = np.arange(S)[None, :] <= batch["length"].reshape(B, 1)
    for key in m.keys:
        diff = np.abs(cut(got[key].cpu().numpy(), case) - g["logits/" + key])
        assert diff[cut(valid, case)].max() <= logit_atol, key
    r = row.cpu().numpy()
    F = len(m.keys)
    assert r[3 * F] == pytest.approx(float(g["data_loss"]), rel=loss_rtol)
    for f, key in enumerate(m.keys):
        assert r[3 * f] == pytest.approx(float(g["loss/" + key]), rel=loss_rtol, abs=1e-5), key
        den = float(g["score_den/" + key])
        assert r[3 * f + 2] == pytest.approx(den, abs=1e-3), key
        assert abs(r[3 * f + 1] - float(g["score_num/" + key])) <= max(1.0, 0.005 * den), key
    # ---- gradients (the engine adds the L2 term inside the optimiser pass: d(l2 sum w^2)/dw = 2 l2 w)
    specs = O.variable_specs(cols, L, 256, input_dtype, CONTEXT.get(case))
    got_grads = eng.get_weights(eng.grads)
    w0 = eng.get_weights()
    total = OrderedDict()
    for name in specs:
        gg = got_grads[name].astype(np.float64).reshape(-1)
        total[name] = gg + 2.0 * L2 * w0[name].astype(np.float64).reshape(-1) if specs[name][2] else gg
    check_gradient_summaries(total, g, grad_tol)
    # ---- L2 + per-variable clipnorm + Adam, one step
    l2_out = torch.zeros(1, device="cuda")
    eng.optimizer_step(1, LR, 1.0, l2_out)
    torch.cuda.synchronize()
    assert float(l2_out.cpu()) == pytest.approx(float(g["total_loss"]) - float(g["data_loss"]), rel=1e-5)
    check_adam_heads({name: w.astype(np.float64).reshape(-1) for name, w in eng.get_weights().items() if name in specs}, g, impl)


# ================================================================================================= demo / eval entry
def _load_demo():
    g = np.load(os.path.join(GOLDEN, "crello_demo.npz"))
    cols = make_input_columns("crello", max_length=50)
    batch = OrderedDict((k[3:], g[k]) for k in g.files if k.startswith("in/"))
    masks = OrderedDict((k[5:], g[k]) for k in g.files if k.startswith("mask/"))
    return g, cols, batch, masks


def test_oracle_demo_call_matches_reference_python():
    """model(example, training=False, demo_args={"masks": ...}) (eval.py:103, notebooks): preprocess_for_test,
    Model.call without dropout, merge_inputs_and_prediction."""
    g, cols, batch, masks = _load_demo()
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    inputs = {k: torch.as_tensor(v) for k, v in batch.items()}
    tmasks = {k: torch.as_tensor(v) for k, v in masks.items()}
    mod = O.preprocess_for_test(inputs, icols, tmasks)
    for key in icols:
        assert np.array_equal(mod[key].numpy(), g["mod/" + key]), key
    params = O.init_params(cols, 2, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05)
    mod64 = {k: (v.double() if v.is_floating_point() else v) for k, v in mod.items()}
    outputs = O.model_forward(params, mod64, icols, 2)
    for key, v in outputs.items():
        assert np.abs(v.numpy() - g["logits/" + key]).max() <= 1e-10, key
    merged = O.merge_inputs_and_prediction(inputs, icols, tmasks, outputs)
    for key in icols:
        ref = g["merged/" + key]
        assert merged[key].shape == ref.shape, key
        assert np.abs(merged[key].numpy() - ref).max() <= 2e-4, key


@pytest.mark.gpu
def test_engine_demo_call_matches_reference_python():
    from flex_dm_b200.mfp import MFP

    g, cols, batch, masks = _load_demo()
    m = MFP(cols, num_blocks=2, masking_method="random", latent_dim=256, dropout=RATE, l2=L2, seed=0)
    params = O.init_params(cols, 2, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05)
    m.set_weights({k: v.numpy().astype(np.float32) for k, v in params.items()})
    out = m(batch, training=False, demo_args={"masks": {k: torch.as_tensor(v) for k, v in masks.items()}})
    torch.cuda.synchronize()
    for f, key in enumerate(m.keys):
        got = m.engine.modified[f].cpu().numpy()
        assert np.array_equal(got, g["mod/" + key].astype(got.dtype)), key
    for key, column in m.input_columns.items():
        ref = g["merged/" + key]
        got = out[key].cpu().numpy()
        assert got.shape == ref.shape, key
        if column["is_sequence"]:
            assert np.abs(got - ref).max() <= H.LOGIT_ATOL, key
        else:
            assert np.array_equal(got, ref), key


# ================================================================================================= iterative decoding
def _load_decode():
    g = np.load(os.path.join(GOLDEN, "crello_decode.npz"))
    cols = make_input_columns("crello", max_length=50)
    batch = OrderedDict((k[3:], g[k]) for k in g.files if k.startswith("in/"))
    masks = OrderedDict((k[5:], g[k]) for k in g.files if k.startswith("mask/"))
    params = O.init_params(cols, 2, 256, WEIGHT_SEED, torch.float64, bias_scale=0.05)
    for name in params:  # decode_weights() of make_golden.py
        if name.startswith("model/decoder/") and name.endswith("/kernel"):
            params[name] = params[name] * 4.0
    return g, cols, batch, masks, params, int(g["num_iter"])


def test_oracle_iterative_decode_matches_reference_python():
    """model(example, training=False, demo_args={"masks": ..., "num_iter": 3}): mfp.py:141-207 run by the reference itself."""
    g, cols, batch, masks, params, num_iter = _load_decode()
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    inputs = {k: (torch.as_tensor(v).double() if v.dtype.kind == "f" else torch.as_tensor(v)) for k, v in batch.items()}
    tmasks = {k: torch.as_tensor(v) for k, v in masks.items()}
    mod = O.preprocess_for_test(inputs, icols, tmasks)
    outputs = O.iterative_decode(params, tmasks, inputs, icols, mod, num_iter, 2)
    merged = O.merge_inputs_and_prediction(inputs, icols, tmasks, outputs)
    for key in icols:
        ref = g["merged/" + key]
        assert merged[key].shape == ref.shape, key
        assert np.abs(merged[key].numpy() - ref).max() <= 1e-9, key


@pytest.mark.gpu
@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-simt", "tf32-tcgen05"])
def test_engine_iterative_decode_matches_reference_python(impl):
    from flex_dm_b200.mfp import MFP

    g, cols, batch, masks, params, num_iter = _load_decode()
    m = MFP(cols, num_blocks=2, masking_method="random", latent_dim=256, dropout=RATE, l2=L2, seed=0)
    m.set_weights({k: v.numpy().astype(np.float32) for k, v in params.items()})
    m.engine.set_gemm_impl(impl)
    out = m(batch, training=False, demo_args={"masks": {k: torch.as_tensor(v) for k, v in masks.items()}, "num_iter": num_iter})
    torch.cuda.synchronize()
    # the decoder kernels are scaled by 4, so are the logits (O(20)): the tolerance is relative to each field's largest logit
    rtol = 1e-4 if impl == 1 else 5e-3
    agree, total = 0, 0
    for key, column in m.input_columns.items():
        ref = g["merged/" + key]
        got = out[key].cpu().numpy()
        assert got.shape == ref.shape, key
        if not column["is_sequence"]:
            assert np.array_equal(got, ref), key
        elif column["type"] == "categorical":
            # which pass a prediction was frozen in depends on a ranking of confidences: compare the decoded labels, and the
            # logits wherever both runs froze the element in the same pass (identical inputs up to rounding)
            atol = rtol * np.abs(ref).max()
            same = np.abs(got - ref).max(axis=(-1, -2)) <= atol
            agree += int((got.argmax(-1) == ref.argmax(-1)).all(axis=-1).sum())
            total += got.shape[0] * got.shape[1]
            assert same.mean() >= 0.9, (key, same.mean(), float(np.abs(got - ref).max()), atol)
        else:
            assert np.abs(got - ref).max() <= rtol * np.abs(ref).max(), (key, float(np.abs(got - ref).max()), float(np.abs(ref).max()))
    assert agree >= 0.97 * total, (agree, total)
