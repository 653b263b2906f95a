"""Reproducibility of the train step (train.py:18-23 seeds everything "for reproducibility and stable validation") and the
single-process semantics of document-sharded data parallelism (train.py:25), on one GPU through the C ABI:

* ``mfp_set_deterministic``: two runs of the same steps give bit-identical weights (split-K weight gradients, bias-gradient column
  sums and LayerNorm gamma / beta gradients summed in a fixed order), and agree with the arrival-order default to rounding;
* ``mfp_set_doc_offset``: a batch cut into shards that carry their global document offset draws the same task ids, masks, random
  tokens and dropout keep-masks as the whole batch, the shards' gradients of ``(1/B_global) * sum`` add up to the whole batch's, and
  the oracle with the same offset reproduces a shard's masks bit for bit."""
import numpy as np
import pytest
import torch

from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from oracle import mfp_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _train(dataset, method, B, S, L, steps, deterministic, seed=3, block_type="deepsvg", impl=0):
    from flex_dm_b200.mfp import MFP, Adam

    cols = make_input_columns(dataset, max_length=S)
    m = MFP(cols, num_blocks=L, block_type=block_type, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=seed)
    m.set_weights(H.perturbed_weights(m.engine, seed))
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    m.set_deterministic(deterministic)
    m.engine.set_gemm_impl(impl)
    batches = [make_synthetic_batch(cols, B, S, seed=10 + i, lengths="ragged") for i in range(2)]
    rows = [m.train_step(batches[i % 2]).clone() for i in range(steps)]
    torch.cuda.synchronize()
    return m.get_weights(), torch.stack(rows).cpu().numpy(), m.engine.get_weights(m.engine.grads)


@pytest.mark.parametrize("dataset,method,B,S,L", [("crello", "random", 64, 128, 2), ("rico", "elem_pos_attr", 9, 20, 1)], ids=["multi-tile", "small"])
def test_deterministic_mode_repeats_bit_for_bit(dataset, method, B, S, L):
    w1, r1, g1 = _train(dataset, method, B, S, L, 4, True)
    w2, r2, g2 = _train(dataset, method, B, S, L, 4, True)
    assert np.array_equal(r1, r2)
    for name in w1:
        assert np.array_equal(w1[name], w2[name]), name
        assert np.array_equal(g1[name], g2[name]), name
    # the arrival-order default computes the same step up to the order of fp32 additions
    w3, r3, g3 = _train(dataset, method, B, S, L, 1, False)
    w4, r4, g4 = _train(dataset, method, B, S, L, 1, True)
    assert r3[0] == pytest.approx(r4[0], rel=1e-5, abs=1e-6)
    for name in g3:
        if name.endswith("dense_key/bias"):
            continue  # exact gradient 0 (softmax is shift-invariant): pure rounding noise of a column sum, in any order
        scale = max(np.abs(g4[name]).max(), 1e-8)
        assert np.abs(g3[name] - g4[name]).max() <= 2e-4 * scale + 1e-9, name


@pytest.mark.parametrize("dataset,method,B,S,L,block_type,impl", [
    ("crello", "random", 256, 128, 1, "deepsvg", 0),      # the benchmarked row count: several tiles per CTA pair (gate words prefetched across tiles)
    ("crello", "elem_pos_attr_img_txt", 64, 128, 2, "deepsvg", 0),   # 64-document shard: narrow tiles on single CTAs
    ("rico", "elem_pos_attr", 9, 20, 2, "transformer", 0),  # post-LayerNorm wiring, a partial row tile
    ("rico", "elem_pos_attr", 9, 20, 1, "deepsvg", 2),      # 3xTF32 GEMMs (the same kernel, three passes)
], ids=["cfg2-rows", "shard64", "postln-small", "3xtf32-small"])
def test_relu_gate_bits_equal_the_mask_operand(monkeypatch, dataset, method, B, S, L, block_type, impl):
    """FFN 1 writes its ReLU gates as one bit per hidden unit and the input-gradient GEMM of FFN 2 reads those words instead of re-reading
    the hidden activation as a mask operand (transformer.py:161-171 through autodiff): a gate is a gate, so steps are bit-identical to the
    mask-operand form (FLEXDM_RELU_BITS=0) in deterministic mode -- metrics rows, gradients and weights."""
    w1, r1, g1 = _train(dataset, method, B, S, L, 2, True, block_type=block_type, impl=impl)
    monkeypatch.setenv("FLEXDM_RELU_BITS", "0")
    w2, r2, g2 = _train(dataset, method, B, S, L, 2, True, block_type=block_type, impl=impl)
    assert np.array_equal(r1, r2)
    for name in w1:
        assert np.array_equal(g1[name], g2[name]), name
        assert np.array_equal(w1[name], w2[name]), name


def _shard_run(cols, method, batch, lo, hi, b_global, S, L, seed, step, impl):
    from flex_dm_b200.mfp import MFP

    m = MFP(cols, num_blocks=L, masking_method=method, latent_dim=256, dropout=0.1, l2=1e-2, seed=0)
    m.set_weights(H.perturbed_weights(m.engine, 0))
    eng = m.engine
    eng.set_gemm_impl(impl)
    eng.set_deterministic(True)
    sub = {k: v[lo:hi] for k, v in batch.items()}
    staged = m.stage(sub)
    _, _, length, dcols = m._bind(staged)
    eng.set_doc_offset(lo)
    tasks = eng.sample_tasks(m.task_ids, seed, step).clone()
    eng.mask_corrupt(length, dcols, tasks, seed, step)
    logits = torch.empty(((hi - lo) * S, eng.logit_width), device="cuda")
    eng.forward(length, None, True, seed, step, logits_out=logits)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / b_global, True, sort_tasks=tasks if m.sort_pos else None)
    eng.backward(length, None, True, seed, step)
    torch.cuda.synchronize()
    return dict(m=m, tasks=tasks.cpu().numpy(), masks=[t.cpu().numpy() for t in eng.masks], mod=[t.cpu().numpy() for t in eng.modified],
                logits=logits.cpu().numpy().reshape(hi - lo, S, -1), row=row.cpu().numpy(), grads=eng.get_weights(eng.grads))


@pytest.mark.parametrize("impl", [1, 0], ids=["fp32-simt", "tf32-tcgen05"])
@pytest.mark.parametrize("dataset,method", [("crello", "random"), ("crello", "elem_pos_attr_img_txt"), ("rico", "elem_pos_attr")])
def test_shards_with_document_offsets_equal_the_whole_batch(dataset, method, impl):
    Bg, S, L, seed, step = 10, 24, 2, 21, 5
    cols = make_input_columns(dataset, max_length=S)
    batch = make_synthetic_batch(cols, Bg, S, seed=2, lengths="ragged")
    whole = _shard_run(cols, method, batch, 0, Bg, Bg, S, L, seed, step, impl)
    bounds = [(0, 4), (4, 7), (7, 10)]  # uneven shards
    parts = [_shard_run(cols, method, batch, lo, hi, Bg, S, L, seed, step, impl) for lo, hi in bounds]
    F = len(whole["masks"])
    for (lo, hi), p in zip(bounds, parts):
        assert np.array_equal(p["tasks"], whole["tasks"][lo:hi])
        for f in range(F):
            assert np.array_equal(p["masks"][f], whole["masks"][f][lo:hi]), f
            assert np.array_equal(p["mod"][f], whole["mod"][f][lo:hi]), f  # random tokens / Gaussian replacements included
        # same rows, same weights, same dropout keep-masks: a row's logits do not depend on which tile of which launch computed it
        valid = np.arange(S)[None, :] <= batch["length"][lo:hi].reshape(-1, 1)
        assert np.allclose(p["logits"][valid], whole["logits"][lo:hi][valid], rtol=0, atol=1e-5 if impl == 1 else 1e-4)
    # additive loss rows and gradients (what the all-reduce sums): shards add up to the whole batch
    row = sum(p["row"] for p in parts)
    assert row[: 3 * F + 1] == pytest.approx(whole["row"][: 3 * F + 1], rel=2e-5, abs=1e-5)
    for name, g in whole["grads"].items():
        if name.endswith("dense_key/bias"):
            continue  # exact gradient 0: rounding noise
        total = sum(p["grads"][name].astype(np.float64) for p in parts)
        scale = max(np.abs(g).max(), 1e-8)
        assert np.abs(total - g).max() <= (2e-4 if impl == 1 else 2e-3) * scale + 1e-9, name
    # ... and the oracle at the same offset draws the shard's masks
    lo, hi = bounds[1]
    m = parts[1]["m"]
    draws = O.PhiloxDraws(seed, step, doc_offset=lo)
    otasks = draws.tasks(hi - lo, m.task_ids)
    assert np.array_equal(otasks, parts[1]["tasks"])
    sub = {k: v[lo:hi] for k, v in batch.items()}
    _, omod, omasks = O.preprocess_for_train(H.to_torch(sub), m.input_columns, torch.as_tensor(otasks), draws)
    for f, key in enumerate(m.keys):
        assert np.array_equal(parts[1]["masks"][f].astype(bool), omasks[key].numpy()), key
