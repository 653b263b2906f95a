#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by executing the REFERENCE'S OWN PYTHON
(/root/reference/src/mfp/mfp: MFP.call, preprocess_for_train, Model.call, LossLayer.call, sort_inputs, ...)
on top of the torch-backed TensorFlow stand-in in oracle/tf_standin/ (TensorFlow cannot be installed here).

    python tests/golden/make_golden.py            # needs /root/reference; runs in the build container only

The vectors travel to the GPU box as committed .npz fixtures; nothing at test / bench time reads /root/reference.

What the vectors pin: the reference's Python logic end to end, at the precision of float64 for the network/loss
and bit-exactly for the masking path.  What they do not pin: the TF/Keras primitive semantics the stand-in restates
from recall and TensorFlow's own RNG streams (random draws are scripted from the B200 path's
Philox contract, in the reference's call order, so that masks are comparable bit for bit).

Per case (file <case>.npz):
  in/<column>            the batch (DataSpec.parse_fn layout)
  tasks                  scripted task ids
  mod/<column>, mask/<column>     outputs of the reference's preprocess_for_train, captured at Model.call / LossLayer.call
  logits/<column>        float64 raw logits of Model.call(modified_inputs, training=True) with the scripted dropout masks
  loss/<column>, score_num/<column>, score_den/<column>, metric/<name>, data_loss, total_loss
  merged/<column>        MFP.call's return value (merge_inputs_and_prediction), float32 run
  gradnorm/<variable>, gradproj/<variable>, gradhead/<variable>   per-variable summaries of d total_loss / d variable
  newproj/<variable>, newhead/<variable>                           the same summaries of the variables after Adam(1e-4, clipnorm=1)
Weights are not stored: they are oracle.mfp_oracle.init_params(cols, L, 256, seed=WEIGHT_SEED, bias_scale=0.05).
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_standin"))
sys.path.insert(1, "/root/reference/src/mfp")
sys.path.insert(2, ROOT)

import tensorflow as tf  # noqa: E402  (the stand-in)
from tensorflow import config as tfc  # noqa: E402
from mfp.models.mfp import MFP as RefMFP  # noqa: E402  (the reference)

from flex_dm_b200.spec import get_valid_input_columns, make_input_columns, make_synthetic_batch  # noqa: E402
from oracle import mfp_oracle as O  # noqa: E402

WEIGHT_SEED = 11
D = 256
RATE = 0.1
L2 = 1e-2
HEAD = 8  # leading entries of each gradient kept verbatim


def projection_vector(name, n):
    """Deterministic +-1 vector per variable (Philox-free, numpy PCG64 keyed by a stable hash of the name)."""
    h = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") ^ (len(name) * 0x9E3779B97F4A7C15 & (2**63 - 1))
    rng = np.random.Generator(np.random.PCG64(h))
    return rng.integers(0, 2, size=n).astype(np.float64) * 2.0 - 1.0


CASES = OrderedDict([
    # name: (dataset, masking_method, B, S, num_blocks, seed, step, lengths, tasks)
    ("crello_random", ("crello", "random", 4, 16, 2, 7, 0, [16, 1, 9, 5], None)),
    ("crello_multi", ("crello", "elem_pos_attr_img_txt", 6, 12, 2, 3, 2, [12, 7, 1, 12, 4, 10], [1, 3, 4, 5, 6, 3])),
    ("rico_pos", ("rico", "elem_pos_attr", 5, 10, 2, 5, 1, [10, 3, 8, 1, 6], [3, 1, 4, 3, 3])),
    # --block_type transformer: the post-LayerNorm TransformerBlock (transformer.py:187-205)
    ("crello_postln", ("crello", "random", 3, 12, 2, 9, 3, [12, 1, 7], None)),
    # --input_dtype shuffled_set: elements shuffled per document (tensor_utils.py:47-76) + PositionEmbedding with dropout (encoder.py:48-55,251-252)
    ("rico_shuffled", ("rico", "elem_pos_attr", 4, 9, 2, 13, 5, [9, 4, 1, 6], [0, 3, 1, 4])),
    # --input_dtype sorted_set: elements sorted by (type, left, top, width, height) (tensor_utils.py:14-44) + PositionEmbedding
    ("crello_sorted", ("crello", "random", 3, 10, 2, 15, 1, [10, 6, 1], None)),
    # --context id / length: a learned special token (task id / document length) joins the sequence (encoder.py:96-110,231-249; decoder.py:74-78).
    # Every document leaves the last row free: the engine keeps the token in the row after a document's last element.
    ("crello_ctx_id", ("crello", "elem_pos_attr_img_txt", 4, 11, 2, 21, 2, [10, 4, 1, 7], [1, 3, 5, 6])),
    ("rico_ctx_length", ("rico", "elem_pos_attr", 4, 9, 2, 30, 1, [8, 3, 1, 5], [3, 1, 4, 3])),
    # --context canvas (token = sum of the canvas columns' embeddings; the decoder gains never-read canvas heads) and canvas_add (that
    # sum added to every element; no token, so the batch may be full length): encoder.py:34-37,177-199,228-249, decoder.py:25-43
    ("crello_ctx_canvas", ("crello", "random", 3, 9, 2, 49, 0, [8, 1, 5], None)),
    ("crello_ctx_canvas_add", ("crello", "elem_pos_attr_img_txt", 3, 8, 2, 27, 1, [8, 1, 6], [4, 1, 6])),
    # --context id with --input_dtype shuffled_set: the token is put in front first and the PositionEmbedding (with its dropout) is added to
    # token + elements afterwards (encoder.py:247-252).  Oracle only: the product path refuses the combination (flex_dm_b200/mfp.py).
    ("rico_ctx_id_shuffled", ("rico", "elem_pos_attr", 4, 9, 2, 37, 2, [8, 3, 1, 5], [0, 3, 1, 4])),
    # ... and --context length with --input_dtype sorted_set on crello (numerical fields, loss conditions)
    ("crello_ctx_length_sorted", ("crello", "random", 3, 10, 2, 33, 1, [9, 4, 1], None)),
])
CONTEXT = {"crello_ctx_id": "id", "rico_ctx_length": "length", "crello_ctx_canvas": "canvas", "crello_ctx_canvas_add": "canvas_add",
           "rico_ctx_id_shuffled": "id", "crello_ctx_length_sorted": "length"}
TOKEN_CONTEXTS = ("id", "length", "canvas")  # contexts that put a token into the sequence
BLOCK_TYPE = {"crello_postln": "transformer"}
INPUT_DTYPE = {"rico_shuffled": "shuffled_set", "crello_sorted": "sorted_set", "rico_ctx_id_shuffled": "shuffled_set", "crello_ctx_length_sorted": "sorted_set"}


def block_dropout_script(draws, B, S, num_blocks, lengths=None):
    """Keep-masks of the two Dropout sites per block, in the engine's row layout ``(B, S, D)``; with a context token (lengths given) the
    engine batch has one padding row more than the reference's (which is cut to the longest document): the masks are gathered into
    the reference's order (token first, oracle.context_dropout_layout) and cut to its S positions (token + S - 1 elements)."""
    keep = {(i, j): torch.from_numpy(draws.dropout_keep(i, j, (B, S, D), RATE)) for i in range(num_blocks) for j in (0, 1)}
    if lengths is not None:
        keep = {k: v[:, :S] for k, v in O.context_dropout_layout(keep, torch.as_tensor(np.asarray(lengths)) - 1).items()}
    return [("dropout", keep[(i, j)].numpy()) for i in range(num_blocks) for j in (0, 1)]


def pos_dropout_script(draws, B, S, lengths=None):
    """Keep-mask of the PositionEmbedding's Dropout (engine row layout); with a context token gathered into the reference's order and
    cut to its positions like the blocks' masks."""
    keep = torch.from_numpy(draws.pos_dropout_keep((B, S, D), RATE))
    if lengths is not None:
        keep = O.context_dropout_layout({"pos": keep}, torch.as_tensor(np.asarray(lengths)) - 1)["pos"][:, :S]
    return [("dropout", keep.numpy())]


def rng_script(cols, B, S, tasks, draws, num_blocks, training, pos_dropout=False, ctx_lengths=None):
    """The reference's RNG call order for one MFP.call(training): mfp.py:301 -> masking.py filter_padding :24-53
    -> random_masking :227-269 -> elem_masking :136-155 -> feat_masking per group :116-133 -> Dropout x2 per block."""
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    seq = get_valid_input_columns(icols)
    fidx = {k: i for i, k in enumerate(seq)}

    def discard(c):
        C = c["shape"][-1]
        return ("randint", None) if c["type"] == "categorical" else ("normal", None)

    Sr = S - 1 if ctx_lengths is not None else S  # columns the reference sees (see run_case)
    s = [("categorical", np.asarray(tasks, dtype=np.int32))]
    for k, c in seq.items():  # filter_padding: apply_token(..., "unused") evaluates the "random" dict entry eagerly
        s.append(discard(c))
    for k, c in seq.items():  # random_masking
        u1, u2, u3 = (u[:, :Sr] for u in draws.uniforms(fidx[k], B, S))
        s += [("uniform", u1), ("uniform", u2), ("uniform", u3), discard(c)]
        C = c["shape"][-1]
        if c["type"] == "categorical":
            s.append(("randint", draws.rand_cat(fidx[k], B, S, C, c["input_dim"])[:, :Sr]))
        else:
            s.append(("normal", draws.rand_num(fidx[k], B, S, C)[:, :Sr]))
    s.append(("uniform", draws.elem_u(B)))  # select_single_element
    for k, c in seq.items():
        s.append(discard(c))
    for group in O.get_attribute_groups(icols.keys()).values():  # feat_masking
        for k in group:
            s.append(discard(seq[k]))
    if training:
        if pos_dropout:  # the encoder's PositionEmbedding dropout comes before the blocks'
            s += pos_dropout_script(draws, B, S, ctx_lengths)
        s += block_dropout_script(draws, B, S, num_blocks, ctx_lengths)
    return s


def set_weights(model, params, dtype):
    variables = model.named_variables()
    assert set(variables) == set(params), (sorted(set(variables) ^ set(params)))
    with torch.no_grad():
        for name, w in variables.items():
            assert tuple(w.shape) == tuple(params[name].shape), (name, w.shape, params[name].shape)
            w.data = params[name].detach().to(dtype).clone()
    return variables


def run_case(name, spec):
    dataset, method, B, S, L, seed, step, lengths, tasks = spec
    cols = make_input_columns(dataset, max_length=max(S, 50))
    batch = make_synthetic_batch(cols, B, S, seed=seed, fixed_lengths=np.asarray(lengths))
    draws = O.PhiloxDraws(seed, step)
    block_type = BLOCK_TYPE.get(name, "deepsvg")
    input_dtype = INPUT_DTYPE.get(name, "set")
    context = CONTEXT.get(name)
    ctx_lengths = lengths if context in TOKEN_CONTEXTS else None
    if input_dtype == "shuffled_set":
        method = method if "random" in method else "random_" + method  # keep task 0 reachable for the scripted task ids
    if input_dtype == "sorted_set":
        icols_ = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
        out_perm = O.sort_inputs({k: torch.as_tensor(v) for k, v in batch.items()}, icols_)[1].numpy().astype(np.int32)
    oracle = O.OracleMFP(cols, num_blocks=L, masking_method=method, dropout=RATE, l2=L2, input_dtype=input_dtype, context=context)
    if tasks is None:
        tasks = draws.tasks(B, oracle.allowed_tasks)
    tasks = np.asarray(tasks, dtype=np.int32)
    params = O.init_params(cols, L, D, WEIGHT_SEED, torch.float64, bias_scale=0.05, input_dtype=input_dtype, context=context)
    out = {"tasks": tasks}
    for k, v in batch.items():
        out["in/" + k] = v
    perm = None
    if input_dtype == "sorted_set":
        out["perm"] = out_perm  # (from the oracle's sort; the sorted columns themselves come from the reference run below)
    if input_dtype == "shuffled_set":
        # the reference shuffles with Python's `random` (tensor_utils.py:60-62); its draws are scripted with the Philox permutation
        perm = draws.shuffle_perm(np.asarray(lengths), S)
        out["perm"] = perm.astype(np.int32)
        import mfp.models.tensor_utils as ref_tu

        class _ScriptedRandom:
            calls = 0

            @staticmethod
            def shuffle(x):
                b = _ScriptedRandom.calls % B
                _ScriptedRandom.calls += 1
                n = len(x)
                x[:] = [int(perm[b, i]) for i in range(n)]

        ref_tu.random = _ScriptedRandom

    # ---------------- phase A: the whole MFP.call in float32 (bit-faithful dtypes for the masking path)
    tfc.FLOAT = torch.float32
    model = RefMFP(cols, num_blocks=L, block_type=block_type, masking_method=method, seq_type="default", arch_type="oneshot",
                   context=context, input_dtype=input_dtype, latent_dim=D, dropout=RATE, l2=L2)
    # with a context token the engine's batch keeps one free row per document; the reference wants S = the longest document
    ref_batch = {k: (v[:, :S - 1] if (context in TOKEN_CONTEXTS and v.ndim == 3) else v) for k, v in batch.items()}
    inputs32 = {k: torch.as_tensor(v).as_subclass(tf.Tensor) for k, v in ref_batch.items()}
    captured = {}
    inner_call, loss_call = model.model.call, model.loss_layer.call

    def model_spy(mod, training):
        captured["mod"] = {k: v.clone() for k, v in mod.items()}
        y = inner_call(mod, training)
        captured["logits"] = {k: v for k, v in y.items()}
        return y

    def loss_spy(inputs, training=False, sort_flag=None, ignore_sort=None):
        captured["targets"], _, captured["masks"] = inputs
        captured["targets"] = dict(captured["targets"])
        captured["masks"] = dict(captured["masks"])
        captured["sort_flag"] = sort_flag
        return loss_call(inputs, training, sort_flag, ignore_sort)

    model.model.call, model.loss_layer.call = model_spy, loss_spy
    tfc.rng = tfc.ScriptedRNG(rng_script(cols, B, S, tasks, draws, L, True, input_dtype != "set", ctx_lengths))
    model({k: v.clone() for k, v in inputs32.items()}, training=True)  # builds the lazily-created variables
    set_weights(model, params, torch.float32)
    model.reset_step_state()
    tfc.rng = tfc.ScriptedRNG(rng_script(cols, B, S, tasks, draws, L, True, input_dtype != "set", ctx_lengths))
    merged = model({k: v.clone() for k, v in inputs32.items()}, training=True)
    assert tfc.rng.done(), "the reference asked for fewer draws than scripted"
    loss32 = float(sum(model.losses))
    assert torch.equal(merged["tasks"].to(torch.int32), torch.as_tensor(tasks))
    for k, v in captured["mod"].items():
        out["mod/" + k] = v.detach().numpy()
    for k, v in captured["masks"].items():
        out["mask/" + k] = v.numpy()
    if input_dtype != "set":
        for k, v in captured["targets"].items():  # the shuffled / sorted batch: what the loss is computed against
            out["tgt/" + k] = v.detach().numpy()
    for k, v in merged.items():
        if k != "tasks":
            out["merged/" + k] = v.detach().numpy().astype(np.float32) if v.is_floating_point() else v.numpy()
    out["total_loss_f32_run"] = np.float64(loss32)

    # ---------------- phase B: Model.call + LossLayer.call in float64 on the captured modified inputs
    tfc.FLOAT = torch.float64
    model.model.call, model.loss_layer.call = inner_call, loss_call
    variables = set_weights(model, params, torch.float64)
    model.reset_step_state()
    mod64 = {k: (v.to(torch.float64) if v.is_floating_point() else v) for k, v in captured["mod"].items()}
    targets64 = {k: (v.to(torch.float64) if v.is_floating_point() else v.clone()) for k, v in captured["targets"].items()}
    tfc.rng = tfc.ScriptedRNG((pos_dropout_script(draws, B, S, ctx_lengths) if input_dtype != "set" else []) +
                              block_dropout_script(draws, B, S, L, ctx_lengths))
    logits = model.model(mod64, True)
    for k, v in logits.items():
        out["logits/" + k] = v.detach().numpy()
    if model.sort_pos:  # mfp.py:335-338
        (scores,) = model.loss_layer((targets64, dict(logits), captured["masks"]), True, torch.as_tensor(tasks) == model.task_names.index("pos"))
    else:
        (scores,) = model.loss_layer((targets64, dict(logits), captured["masks"]), True)
    assert tfc.rng.done()
    data_loss = sum(model.loss_layer._losses)
    total = sum(model.losses)
    for k, v in scores.items():
        key, what = k.rsplit("_score_", 1)
        out["score_%s/%s" % (what, key)] = np.float64(v.detach())
    for mname, v in model.loss_layer.step_metrics:
        if mname.endswith("_loss"):
            out["loss/" + mname[:-5]] = np.float64(torch.as_tensor(v).detach())
        else:
            out["metric/" + mname] = np.float64(torch.as_tensor(v).detach())
    out["data_loss"] = np.float64(data_loss.detach())
    out["total_loss"] = np.float64(total.detach())
    assert abs(loss32 - float(total)) <= 2e-4 * abs(float(total)), (loss32, float(total))

    names = list(variables.keys())
    grads = torch.autograd.grad(total, [variables[n] for n in names], allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(variables[n]) for g, n in zip(grads, names)]
    opt = tf.keras.optimizers.Adam(learning_rate=1e-4, clipnorm=1.0)  # train.py:71-77
    opt.apply_gradients(zip(grads, [variables[n] for n in names]))
    for n, g in zip(names, grads):
        g = g.detach().numpy().reshape(-1)
        w = variables[n].detach().numpy().reshape(-1)
        pv = projection_vector(n, g.size)
        out["gradnorm/" + n] = np.float64(np.linalg.norm(g))
        out["gradproj/" + n] = np.float64(g @ pv)
        out["gradhead/" + n] = g[:HEAD].copy()
        out["newproj/" + n] = np.float64(w @ pv)
        out["newhead/" + n] = w[:HEAD].copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-16s tasks=%s total_loss=%.6f data_loss=%.6f -> %s (%.0f kB)" % (name, tasks.tolist(), float(total), float(data_loss),
                                                                            os.path.relpath(path, ROOT), os.path.getsize(path) / 1e3))


def run_demo_case(name="crello_demo", B=3, S=10, L=2, seed=5, lengths=(10, 1, 6), masked_keys=("left", "color", "image_embedding")):
    """The eval.py / notebook entry: model(example, training=False, demo_args={"masks": ...}) (mfp.py:298-347 with
    preprocess_for_test :72-92); forward only, no loss layer."""
    cols = make_input_columns("crello", max_length=50)
    batch = make_synthetic_batch(cols, B, S, seed=seed, fixed_lengths=np.asarray(lengths))
    params = O.init_params(cols, L, D, WEIGHT_SEED, torch.float64, bias_scale=0.05)
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    seq = get_valid_input_columns(icols)
    seq_mask = np.arange(S)[None, :] < np.asarray(lengths)[:, None]
    masks = {}
    for k, c in icols.items():
        masks[k] = np.ones((B,), bool) if not c["is_sequence"] else (seq_mask.copy() if k in masked_keys else np.zeros((B, S), bool))
    out = {"tasks": np.zeros((B,), np.int32)}
    for k, v in batch.items():
        out["in/" + k] = v
    for k, v in masks.items():
        out["mask/" + k] = v

    def script():
        d = [("randint", None) if c["type"] == "categorical" else ("normal", None) for c in seq.values()]
        return [("categorical", out["tasks"])] + d + d  # task sampler, filter_padding, apply_token(..., "masked")

    tfc.FLOAT = torch.float32
    model = RefMFP(cols, num_blocks=L, block_type="deepsvg", masking_method="random", seq_type="default", arch_type="oneshot",
                   context=None, input_dtype="set", latent_dim=D, dropout=RATE, l2=L2)
    captured = {}
    inner_call = model.model.call

    def model_spy(mod, training):
        assert not training
        captured["mod"] = {k: v.clone() for k, v in mod.items()}
        return inner_call(mod, training)

    model.model.call = model_spy
    inputs32 = {k: torch.as_tensor(v).as_subclass(tf.Tensor) for k, v in batch.items()}
    tmasks = {k: torch.as_tensor(v).as_subclass(tf.Tensor) for k, v in masks.items()}
    tfc.rng = tfc.ScriptedRNG(script())
    model(dict(inputs32), training=False, demo_args={"masks": dict(tmasks)})
    set_weights(model, params, torch.float32)
    tfc.rng = tfc.ScriptedRNG(script())
    merged = model(dict(inputs32), training=False, demo_args={"masks": dict(tmasks)})
    assert tfc.rng.done()
    for k, v in captured["mod"].items():
        out["mod/" + k] = v.detach().numpy()
    for k, v in merged.items():
        if k != "tasks":
            out["merged/" + k] = v.detach().numpy().astype(np.float32) if v.is_floating_point() else v.numpy()
    tfc.FLOAT = torch.float64
    model.model.call = inner_call
    set_weights(model, params, torch.float64)
    mod64 = {k: (v.to(torch.float64) if v.is_floating_point() else v) for k, v in captured["mod"].items()}
    tfc.rng = tfc.ScriptedRNG([])
    for k, v in model.model(mod64, False).items():
        out["logits/" + k] = v.detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-16s demo forward -> %s (%.0f kB)" % (name, os.path.relpath(path, ROOT), os.path.getsize(path) / 1e3))


def decode_weights(cols, L):
    """Weights of the iterative-decoding case: decoder kernels scaled (x4) so that the per-field confidences (max softmax
    probability) are spread out without saturating at 1.0 (where float32 ties would make the top-k selection arbitrary)."""
    params = O.init_params(cols, L, D, WEIGHT_SEED, torch.float64, bias_scale=0.05)
    for name in params:
        if name.startswith("model/decoder/") and name.endswith("/kernel"):
            params[name] = params[name] * 4.0
    return params


def run_decode_case(name="crello_decode", B=1, S=12, L=2, seed=6, lengths=(12,), num_iter=3,
                    masked_keys=("type", "left", "top", "width", "height", "color", "image_embedding")):
    """eval.py --num_iter > 1: MaskGIT-like iterative decoding (mfp.py:141-207), whole MFP.call in float64.  B = 1: the reference
    compares a (B, S) confidence with a (B,) threshold (mfp.py:184), which only broadcasts for a single document."""
    cols = make_input_columns("crello", max_length=50)
    batch = make_synthetic_batch(cols, B, S, seed=seed, fixed_lengths=np.asarray(lengths))
    params = decode_weights(cols, L)
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    seq = get_valid_input_columns(icols)
    seq_mask = np.arange(S)[None, :] < np.asarray(lengths)[:, None]
    masks = {}
    for k, c in icols.items():
        masks[k] = np.ones((B,), bool) if not c["is_sequence"] else (seq_mask.copy() if k in masked_keys else np.zeros((B, S), bool))
    out = {"tasks": np.zeros((B,), np.int32), "num_iter": np.int32(num_iter)}
    for k, v in batch.items():
        out["in/" + k] = v
    for k, v in masks.items():
        out["mask/" + k] = v
    d = [("randint", None) if c["type"] == "categorical" else ("normal", None) for c in seq.values()]
    # task sampler; preprocess_for_test: filter_padding + apply_token; iterative_decode: filter_padding, then apply_token per iteration
    script = [("categorical", out["tasks"])] + d + d + d + d * num_iter
    tfc.FLOAT = torch.float64
    model = RefMFP(cols, num_blocks=L, block_type="deepsvg", masking_method="random", seq_type="default", arch_type="oneshot",
                   context=None, input_dtype="set", latent_dim=D, dropout=RATE, l2=L2)

    def inputs64():
        return {k: (torch.as_tensor(v).double() if v.dtype.kind == "f" else torch.as_tensor(v)).as_subclass(tf.Tensor) for k, v in batch.items()}

    tmasks = {k: torch.as_tensor(v).as_subclass(tf.Tensor) for k, v in masks.items()}
    tfc.rng = tfc.ScriptedRNG(script)
    model(inputs64(), training=False, demo_args={"masks": dict(tmasks), "num_iter": num_iter})
    set_weights(model, params, torch.float64)
    tfc.rng = tfc.ScriptedRNG(script)
    merged = model(inputs64(), training=False, demo_args={"masks": dict(tmasks), "num_iter": num_iter})
    assert tfc.rng.done()
    for k, v in merged.items():
        if k != "tasks":
            out["merged/" + k] = v.detach().numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-16s iterative decode (%d iterations) -> %s (%.0f kB)" % (name, num_iter, os.path.relpath(path, ROOT), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for case, spec in CASES.items():
        if not only or case in only:
            run_case(case, spec)
    if not only or "crello_demo" in only:
        run_demo_case()
    if not only or "crello_decode" in only:
        run_decode_case()
