"""GPU parity tests: the CUDA engine (through the C ABI, via flex_dm_b200.engine / MFP) against the CPU oracle
on identical seeded inputs and weights.  Integer / mask work is compared bit-exactly; floating point within the
tolerances stated in tests/helpers.py."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from oracle import mfp_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

CONFIGS = [
    ("crello", "random", 4, 32, 2),                   # BASELINE.json configs[0] (cfg1)
    ("crello", "elem_pos_attr_img_txt", 5, 20, 2),
    ("rico", "elem_pos_attr", 6, 24, 2),
]


def _model(dataset, method, num_blocks, dropout=0.0, l2=1e-2, seed=0):
    from flex_dm_b200.mfp import MFP

    cols = make_input_columns(dataset)
    m = MFP(cols, num_blocks=num_blocks, masking_method=method, latent_dim=256, dropout=dropout, l2=l2, seed=seed)
    m.set_weights(H.perturbed_weights(m.engine, seed))
    return cols, m


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1)], ids=["fwd", "dgrad", "wgrad"])
@pytest.mark.parametrize("M,N,K,splits", [(128, 128, 32, 1), (256, 256, 256, 1), (200, 1380, 256, 1), (256, 512, 1000, 4), (132, 36, 72, 1)])
def test_gemm_layouts(impl, a_mn, b_mn, M, N, K, splits):
    from flex_dm_b200.engine import debug_gemm

    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g)
    Bm = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    ref = A.double() @ Bm.double().T + bias.double()
    Ad = (A.T.contiguous() if a_mn else A.contiguous()).cuda()
    Bd = (Bm.T.contiguous() if b_mn else Bm.contiguous()).cuda()
    out = torch.zeros(M, N, device="cuda")
    debug_gemm(Ad, a_mn, Bd, b_mn, M, N, K, bias=bias.cuda(), splits=splits, impl=impl, out=out)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    scale = (A.double().abs() @ Bm.double().abs().T).max().item()
    tol = 1e-5 if impl == 1 else 2e-3
    assert err <= tol * scale, (err, scale)


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (1000, 520, 64), (128, 36, 40)])
def test_gemm_fused_epilogues(impl, M, N, K):
    """bias + ReLU, residual add (in place), ReLU mask of a dgrad, and the fused bias-gradient column sum of a wgrad."""
    from flex_dm_b200.engine import debug_gemm

    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g)
    bias = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g)
    tol = 1e-5 if impl == 1 else 2e-3
    scale = (A.double().abs() @ W.double().abs().T).max().item()
    Ad, Wd, bd = A.cuda(), W.cuda(), bias.cuda()
    # bias + relu
    out = debug_gemm(Ad, 0, Wd, 0, M, N, K, bias=bd, relu=True, impl=impl)
    ref = torch.relu(A.double() @ W.double().T + bias.double())
    assert (out.cpu().double() - ref).abs().max().item() <= tol * scale
    # residual, in place (out aliases the residual, as in the encoder)
    x = R.cuda().clone()
    debug_gemm(Ad, 0, Wd, 0, M, N, K, bias=bd, impl=impl, out=x, residual=x)
    ref = A.double() @ W.double().T + bias.double() + R.double()
    assert (x.cpu().double() - ref).abs().max().item() <= tol * scale
    # relu mask
    out = debug_gemm(Ad, 0, Wd, 0, M, N, K, impl=impl, relu_src=R.cuda())
    ref = (A.double() @ W.double().T) * (R.double() > 0)
    assert (out.cpu().double() - ref).abs().max().item() <= tol * scale
    # wgrad with split-K and the fused column sum: D[K', N'] = X^T . dY, colsum = sum_rows dY   (rows = M tokens)
    X = torch.randn(M, K, generator=g)
    dY = torch.randn(M, N, generator=g)
    out = torch.zeros(K, N, device="cuda")
    cs = torch.zeros(N, device="cuda")
    debug_gemm(X.cuda(), 1, dY.cuda(), 1, K, N, M, splits=3, impl=impl, out=out, colsum=cs)
    ref = X.double().T @ dY.double()
    wscale = (X.double().abs().T @ dY.double().abs()).max().item()
    assert (out.cpu().double() - ref).abs().max().item() <= tol * wscale
    assert (cs.cpu().double() - dY.double().sum(0)).abs().max().item() <= 1e-4 * dY.abs().sum(0).max().item()


def test_gemm_cta_pair_tiles():
    """Problems large enough for the 256-row tiles of CTA pairs (tcgen05 cta_group::2: each CTA stages its 128 rows of A and half of the
    B tile): forward layout with bias + ReLU and a ragged last tile, residual in place, dgrad with a ReLU mask, and a split-K weight
    gradient with the fused column sum (the peer CTA sums its own half of the columns)."""
    from flex_dm_b200.engine import debug_gemm

    M, N, K = 6400 + 72, 512, 256
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(K, N, generator=g)   # forward: B is [K][N] (MN-major)
    bias = torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g)
    scale = (A.double().abs() @ W.double().abs()).max().item()
    out = debug_gemm(A.cuda(), 0, W.cuda(), 1, M, N, K, bias=bias.cuda(), relu=True, impl=0)
    ref = torch.relu(A.double() @ W.double() + bias.double())
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-3 * scale
    x = R.cuda().clone()
    debug_gemm(A.cuda(), 0, W.cuda(), 1, M, N, K, bias=bias.cuda(), impl=0, out=x, residual=x)
    ref = A.double() @ W.double() + bias.double() + R.double()
    assert (x.cpu().double() - ref).abs().max().item() <= 2e-3 * scale
    Wk = torch.randn(N, K, generator=g)  # dgrad: B is [N][K] (K-major)
    out = debug_gemm(A.cuda(), 0, Wk.cuda(), 0, M, N, K, impl=0, relu_src=R.cuda())
    ref = (A.double() @ Wk.double().T) * (R.double() > 0)
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-3 * scale
    # weight gradient: D[512, 256] = X^T . dY over M tokens, split-K, colsum = sum_rows dY
    X = torch.randn(M, 512, generator=g)
    dY = torch.randn(M, 256, generator=g)
    out = torch.zeros(512, 256, device="cuda")
    cs = torch.zeros(256, device="cuda")
    debug_gemm(X.cuda(), 1, dY.cuda(), 1, 512, 256, M, splits=37, impl=0, out=out, colsum=cs)
    ref = X.double().T @ dY.double()
    wscale = (X.double().abs().T @ dY.double().abs()).max().item()
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-3 * wscale
    assert (cs.cpu().double() - dY.double().sum(0)).abs().max().item() <= 1e-4 * dY.abs().sum(0).max().item()


# ----------------------------------------------------------------------------------------------- attention core
def _attention_reference(qkv, length, B, S):
    """transformer.py:60-76 in float64: softmax(QK^T/sqrt(dh) - 1e9 (1 - key mask)) V per head."""
    x = qkv.double().reshape(B, S, 3, 8, 32)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))  # (B, H, S, dh)
    score = q @ k.transpose(-1, -2) / np.sqrt(32.0)
    mask = (torch.arange(S)[None, :] <= length[:, None]).double()  # (B, S)
    score = score + (-1e9) * (1.0 - mask)[:, None, None, :]
    lse = torch.logsumexp(score, dim=-1)
    out = torch.softmax(score, dim=-1) @ v
    return out.permute(0, 2, 1, 3).reshape(B * S, 256), lse


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("B,S", [(3, 128), (5, 50), (2, 7), (40, 128)])
def test_attention_core(impl, B, S):
    from flex_dm_b200.engine import debug_attention

    g = torch.Generator().manual_seed(B * 1000 + S)
    qkv = torch.randn(B * S, 768, generator=g) * 1.5
    length = torch.randint(0, S, (B,), generator=g, dtype=torch.int32)
    length[0] = S - 1
    if B > 1:
        length[1] = 0  # a single valid element: every query returns V[0]
    ref_out, ref_lse = _attention_reference(qkv, length, B, S)
    out, lse = debug_attention(qkv.cuda(), length.cuda(), B, S, impl=impl)
    torch.cuda.synchronize()
    tol = 2e-5 if impl == 1 else 4e-3
    assert (out.cpu().double() - ref_out).abs().max().item() <= tol * max(1.0, ref_out.abs().max().item())
    assert (lse.cpu().double() - ref_lse).abs().max().item() <= (1e-4 if impl == 1 else 2e-2)


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("B,S", [(3, 128), (5, 50), (2, 7), (40, 128)])
def test_attention_core_backward(impl, B, S):
    from flex_dm_b200.engine import debug_attention, debug_attention_bwd

    g = torch.Generator().manual_seed(B * 77 + S)
    qkv = torch.randn(B * S, 768, generator=g) * 1.2
    length = torch.randint(0, S, (B,), generator=g, dtype=torch.int32)
    length[0] = S - 1
    valid = (torch.arange(S)[None, :] <= length[:, None]).reshape(B * S, 1)
    dout = torch.randn(B * S, 256, generator=g) * valid  # padded elements carry no upstream gradient (metrics.py:263-267)
    x = qkv.double().clone().requires_grad_(True)
    ref_out, _ = _attention_reference(x, length, B, S)
    (ref_out * dout.double()).sum().backward()
    ref = x.grad
    out, lse = debug_attention(qkv.cuda(), length.cuda(), B, S, impl=impl)
    dqkv = debug_attention_bwd(qkv.cuda(), length.cuda(), B, S, out, lse, dout.cuda(), impl=impl)
    torch.cuda.synchronize()
    got = dqkv.cpu().double()
    assert torch.isfinite(got).all()
    for name, sl in (("dq", slice(0, 256)), ("dk", slice(256, 512)), ("dv", slice(512, 768))):
        err = (got[:, sl] - ref[:, sl]).norm().item() / max(ref[:, sl].norm().item(), 1e-30)
        assert err <= (1e-5 if impl == 1 else 3e-3), (name, err)


# ----------------------------------------------------------------------------------------------- masking (bit-exact)
@pytest.mark.parametrize("dataset,method,B,S,L", CONFIGS)
def test_mask_corrupt_matches_oracle(dataset, method, B, S, L):
    cols, m = _model(dataset, method, 1)
    batch = make_synthetic_batch(cols, B, S, seed=3, lengths="ragged")
    staged = m.stage(batch)
    Bq, Sq, length, dcols = m._bind(staged)
    for step in range(3):
        tasks = m.engine.sample_tasks(m.task_ids, 11, step).clone()
        m.engine.mask_corrupt(length, dcols, tasks, 11, step)
        torch.cuda.synchronize()
        draws = O.PhiloxDraws(11, step)
        ot = draws.tasks(B, m.task_ids)
        assert np.array_equal(tasks.cpu().numpy(), ot)
        _, omod, omasks = H.oracle_train_inputs(cols, batch, ot, 11, step)
        for f, key in enumerate(m.keys):
            got = m.engine.modified[f].cpu()
            assert np.array_equal(m.engine.masks[f].cpu().numpy().astype(bool), omasks[key].numpy()), key
            if cols[key]["type"] == "categorical":
                assert np.array_equal(got.numpy(), omod[key].numpy().astype(np.int32)), key
            else:
                assert np.allclose(got.numpy(), omod[key].numpy(), atol=1e-6, rtol=0), key


def test_mask_for_test_matches_oracle():
    cols, m = _model("crello", "random", 1)
    B, S = 3, 10
    batch = make_synthetic_batch(cols, B, S, seed=5, lengths="ragged")
    t = H.to_torch(batch)
    seq = O.get_seq_mask(t["length"], S)
    masks = O.get_initial_masks(m.input_columns, seq)
    for key in ("left", "color", "image_embedding"):
        masks[key] = seq
    omod = O.preprocess_for_test(t, m.input_columns, masks)
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    m.engine.mask_for_test(length, dcols, [masks[k].to(torch.uint8).cuda() for k in m.keys])
    for f, key in enumerate(m.keys):
        assert np.allclose(m.engine.modified[f].cpu().numpy(), omod[key].numpy(), atol=0), key


# ----------------------------------------------------------------------------------------------- forward / loss / backward
def _oracle_run(cols, m, batch, tasks, seed, step, drop=None):
    targets, omod, omasks = H.oracle_train_inputs(cols, batch, tasks, seed, step)
    params = OrderedDict((k, v.requires_grad_(True)) for k, v in H.oracle_params_from_engine(m.engine).items())
    L = m.engine.cfg.num_blocks
    outputs = O.model_forward(params, omod, m.input_columns, L, drop, float(m.engine.cfg.dropout))
    sort_flag = (torch.as_tensor(tasks) == m.task_names.index("pos")) if m.sort_pos else None
    total, losses, scores, metrics = O.loss_layer(targets, outputs, omasks, cols, sort_flag)
    total.backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()) for k, v in params.items())
    return outputs, float(total.detach()), losses, scores, metrics, grads, omod, omasks


@pytest.mark.parametrize("impl", [1, 0, 2], ids=["fp32-simt", "tf32-tcgen05", "3xtf32-tcgen05"])
@pytest.mark.parametrize("dataset,method,B,S,L", CONFIGS)
def test_forward_loss_backward_match_oracle(dataset, method, B, S, L, impl):
    cols, m = _model(dataset, method, L)
    m.engine.set_gemm_impl(impl)
    logit_atol, loss_rtol, grad_tol = H.tolerances(impl)
    batch = make_synthetic_batch(cols, B, S, seed=1, lengths="ragged")
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    eng = m.engine
    seed, step = 5, 0
    tasks = eng.sample_tasks(m.task_ids, seed, step).clone()
    eng.mask_corrupt(length, dcols, tasks, seed, step)
    logits = torch.empty((B * S, eng.logit_width), device="cuda")
    eng.forward(length, None, False, seed, step, logits_out=logits)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True, sort_tasks=tasks if m.sort_pos else None)
    eng.backward(length, None, False, seed, step)
    torch.cuda.synchronize()

    outputs, total, losses, scores, metrics, grads, _, _ = _oracle_run(cols, m, batch, tasks.cpu().numpy(), seed, step)
    got = m.split_logits(logits, B, S)
    for key in m.keys:
        ref = outputs[key].detach().numpy()
        err = np.abs(got[key].cpu().numpy() - ref).max()
        assert err <= logit_atol and err <= H.LOGIT_RTOL * max(1.0, np.abs(ref).max()) * 4, (key, err)
    r = row.cpu().numpy()
    F = len(m.keys)
    assert r[3 * F] == pytest.approx(total, rel=loss_rtol)
    for f, key in enumerate(m.keys):
        assert r[3 * f] == pytest.approx(float(losses[key]), rel=loss_rtol, abs=1e-5), key
        assert r[3 * f + 2] == pytest.approx(float(scores[key + "_score_den"]), abs=1e-3), key
        # the score numerator counts argmax hits: allow one flip per 200 from TF32 near-ties
        assert abs(r[3 * f + 1] - float(scores[key + "_score_num"])) <= max(1.0, 0.005 * float(scores[key + "_score_den"])), key
    got_grads = eng.get_weights(eng.grads)
    for name, g in grads.items():
        gn = np.linalg.norm(g)
        if gn < 1e-9:
            # e.g. the key bias: softmax is shift-invariant, so its exact gradient is 0 and what is left is the rounding
            # noise of the column sum of dK -- bound it relative to the gradient of the kernel it belongs to
            sibling = np.linalg.norm(grads[name.replace("/bias", "/kernel")])
            assert np.linalg.norm(got_grads[name]) <= grad_tol * max(sibling, 1e-6), name
        else:
            assert H.rel_l2(got_grads[name], g) <= grad_tol, (name, H.rel_l2(got_grads[name], g))


def test_training_dropout_matches_oracle_keep_masks():
    cols, m = _model("crello", "random", 2, dropout=0.1)
    B, S = 4, 16
    batch = make_synthetic_batch(cols, B, S, seed=2, lengths="ragged")
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    eng = m.engine
    seed, step = 9, 4
    tasks = eng.sample_tasks(m.task_ids, seed, step).clone()
    eng.mask_corrupt(length, dcols, tasks, seed, step)
    logits = torch.empty((B * S, eng.logit_width), device="cuda")
    eng.forward(length, None, True, seed, step, logits_out=logits)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
    eng.backward(length, None, True, seed, step)
    torch.cuda.synchronize()
    draws = O.PhiloxDraws(seed, step)
    drop = {(i, j): torch.from_numpy(draws.dropout_keep(i, j, (B, S, 256), 0.1)) for i in range(2) for j in (0, 1)}
    outputs, total, losses, scores, metrics, grads, _, _ = _oracle_run(cols, m, batch, tasks.cpu().numpy(), seed, step, drop)
    got = m.split_logits(logits, B, S)
    for key in m.keys:
        assert np.abs(got[key].cpu().numpy() - outputs[key].detach().numpy()).max() <= H.LOGIT_ATOL, key
    assert row.cpu().numpy()[3 * len(m.keys)] == pytest.approx(total, rel=H.LOSS_RTOL)
    got_grads = eng.get_weights(eng.grads)
    for name in ("model/blocks/seq2seq/seq2seq_0/attn/dense_query/kernel", "model/blocks/seq2seq/seq2seq_1/mlp/layer_with_weights-1/kernel",
                 "model/encoder/input_layer/left/embeddings"):
        assert H.rel_l2(got_grads[name], grads[name]) <= H.GRAD_REL_L2, name


def test_backward_stages_equal_the_monolithic_pass():
    """mfp_backward_stages (what the data-parallel step runs, one all-reduce slice per stage) == mfp_backward; the stage ranges
    tile the flat gradient buffer."""
    cols, m = _model("crello", "random", 2, dropout=0.1)
    B, S = 4, 16
    batch = make_synthetic_batch(cols, B, S, seed=2, lengths="ragged")
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    eng = m.engine
    seed, step = 9, 4
    tasks = eng.sample_tasks(m.task_ids, seed, step).clone()
    eng.mask_corrupt(length, dcols, tasks, seed, step)
    eng.forward(length, None, True, seed, step)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
    eng.backward(length, None, True, seed, step)
    ref = eng.grads.clone()
    ranges = eng.backward_stage_ranges()
    assert len(ranges) == 2 + 2
    spans = sorted(ranges)
    assert spans[0][0] == 0 and spans[-1][1] == eng.param_count and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    eng.grads.fill_(float("nan"))
    final = torch.zeros_like(ref, dtype=torch.bool)
    for s, (lo, hi) in enumerate(ranges):
        eng.backward_stages(length, s, s, None, True, seed, step)
        torch.cuda.synchronize()
        final[lo:hi] = True
        got = eng.grads[final]  # everything declared final so far already equals the monolithic result (split-K summation order aside)
        assert torch.isfinite(got).all()
        assert float((got - ref[final]).abs().max()) <= 1e-5 * float(ref.abs().max()), s
    assert bool(final.all())


def test_optimizer_step_matches_oracle():
    cols, m = _model("rico", "random", 1)
    eng = m.engine
    eng.bind(2, 8)
    rng = np.random.Generator(np.random.PCG64(0))
    names = list(eng.variables.keys())
    grads = OrderedDict((n, (rng.standard_normal(eng.variable_shape(n)) * (3.0 if i % 3 == 0 else 0.01)).astype(np.float32)) for i, n in enumerate(names))
    p = H.oracle_params_from_engine(eng)
    specs = O.variable_specs(cols, 1, 256)
    om = OrderedDict((k, torch.zeros_like(v)) for k, v in p.items())
    ov = OrderedDict((k, torch.zeros_like(v)) for k, v in p.items())
    l2 = float(eng.cfg.l2)
    l2_out = torch.zeros(1, device="cuda")
    for t in (1, 2, 3):
        eng.grads.zero_()
        for n in names:
            eng.variable_view(n, eng.grads).copy_(torch.from_numpy(grads[n]).reshape(eng.variable_view(n).shape).cuda())
        expect_l2 = float(O.l2_regulariser(p, specs, l2))
        og = OrderedDict((n, torch.tensor(grads[n], dtype=torch.float64) + (2 * l2 * p[n] if specs[n][2] else 0.0)) for n in names)
        O.adam_step(p, og, om, ov, t, lr=1e-3, clipnorm=1.0)
        eng.optimizer_step(t, 1e-3, 1.0, l2_out)
        torch.cuda.synchronize()
        assert float(l2_out.cpu()) == pytest.approx(expect_l2, rel=1e-5)
    got = eng.get_weights()
    for n in names:
        assert np.abs(got[n] - p[n].numpy()).max() <= H.WEIGHT_ATOL * 3, n
    # padding between variables is never touched
    used = torch.zeros(eng.param_count, dtype=torch.bool)
    for n in names:
        eng.variable_view(n, used).fill_(True)
    assert float(eng.params.cpu()[~used].abs().max()) == 0.0


def test_train_steps_track_oracle_loss():
    """'loss match' (BASELINE.json metric): same weights, same Philox-defined masks and dropout -> same loss curve."""
    cols, m = _model("crello", "random", 2, dropout=0.1, l2=1e-2, seed=3)
    from flex_dm_b200.mfp import Adam

    m.seed = 21
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    o = O.OracleMFP(cols, num_blocks=2, masking_method="random", dropout=0.1, l2=1e-2, dtype=torch.float64, learning_rate=1e-3, clipnorm=1.0)
    o.params = H.oracle_params_from_engine(m.engine)
    o.m = OrderedDict((k, torch.zeros_like(v)) for k, v in o.params.items())
    o.v = OrderedDict((k, torch.zeros_like(v)) for k, v in o.params.items())
    batch = make_synthetic_batch(cols, 4, 32, seed=0, lengths="ragged")
    w0 = m.get_weights()
    for step in range(4):
        row = m.train_step(batch)
        got = m.metrics_from_row(row)
        ref = o.train_step(batch, seed=21, step=step)
        # after two Adam updates of lr = 1e-3 the two runs no longer hold the same weights: Adam normalises every gradient entry to ~lr,
        # so TF32-sized errors on near-zero entries become lr-sized weight differences (see the weight checks below)
        assert got["loss"] == pytest.approx(ref["loss"], rel=H.LOSS_RTOL if step < 2 else 5 * H.LOSS_RTOL), step
        assert got["total_score"] == pytest.approx(ref["metrics"]["total_score"], abs=2e-2)
    w = m.get_weights()
    for name in ("model/blocks/seq2seq/seq2seq_1/attn/combine_heads/kernel", "model/decoder/decoders/left/kernel"):
        # Adam moves every weight by ~lr per step whatever |g| is, so a TF32-sized gradient error on a near-zero
        # gradient entry can flip a whole lr-sized update: compare the accumulated update in relative L2, and bound
        # the worst entry by the 4 * lr a weight can travel in 4 steps.
        delta_ref = o.params[name].numpy() - w0[name]
        assert H.rel_l2(w[name] - w0[name], delta_ref) < 0.1, name
        assert np.abs(w[name] - o.params[name].numpy()).max() < 4 * 1e-3, name


def test_call_outputs_merge_ground_truth():
    cols, m = _model("crello", "random", 1)
    B, S = 3, 12
    batch = make_synthetic_batch(cols, B, S, seed=8, lengths="ragged")
    t = H.to_torch(batch)
    seq = O.get_seq_mask(t["length"], S)
    masks = O.get_initial_masks(m.input_columns, seq)
    masks["top"] = seq
    masks["text_embedding"] = seq
    out = m(batch, training=False, demo_args={"masks": masks})
    omod = O.preprocess_for_test(t, m.input_columns, masks)
    ref = O.model_forward(H.oracle_params_from_engine(m.engine), omod, m.input_columns, 1)
    merged = O.merge_inputs_and_prediction(t, m.input_columns, masks, ref)
    for key in m.keys:
        assert out[key].shape == merged[key].shape, key
        assert np.abs(out[key].cpu().numpy() - merged[key].numpy()).max() <= H.LOGIT_ATOL, key
    assert np.array_equal(out["left"].cpu().numpy(), np.eye(64, dtype=np.float32)[batch["left"]])  # unmasked -> one-hot GT
    assert np.array_equal(out["canvas_width"].cpu().numpy(), batch["canvas_width"])
    assert out["tasks"].shape == (B,)


def test_loss_layer_standalone_with_sort():
    from flex_dm_b200.metrics import LossLayer

    cols = make_input_columns("rico")
    B, S = 4, 9
    batch = make_synthetic_batch(cols, B, S, seed=4, lengths="ragged")
    t = H.to_torch(batch)
    seq = O.get_seq_mask(t["length"], S)
    icols = OrderedDict((k, v) for k, v in cols.items())
    masks = O.get_initial_masks(icols, seq)
    for key in ("left", "top", "width", "height"):
        masks[key] = seq
    g = torch.Generator().manual_seed(0)
    pred = OrderedDict()
    for key, c in O.get_valid_input_columns(cols).items():
        pred[key] = torch.randn(B, S, c["shape"][-1], c["input_dim"], generator=g)
    flag = torch.tensor([True, False, True, True])
    _, losses, scores, _ = O.loss_layer(t, {k: v.double() for k, v in pred.items()}, masks, cols, flag)
    layer = LossLayer(cols)
    (got,) = layer((batch, pred, masks), False, flag)
    for k, v in scores.items():
        assert float(got[k]) == pytest.approx(float(v), abs=1e-4), k
    for key in layer.keys:
        assert layer.losses[key] == pytest.approx(float(losses[key]), rel=1e-4, abs=1e-5), key


# ----------------------------------------------------------------------------------------------- full-size properties
def test_full_size_step_properties():
    """BASELINE.json configs[1] shape (B=256, S=128, L=4): size-independent properties instead of the oracle."""
    from flex_dm_b200.mfp import Adam

    cols, m = _model("crello", "random", 4, dropout=0.0, l2=1e-2)
    B, S = 256, 128
    batch = make_synthetic_batch(cols, B, S, seed=0, lengths="ragged")
    staged = m.stage(batch)
    _, _, length, dcols = m._bind(staged)
    eng = m.engine
    tasks = eng.sample_tasks(m.task_ids, 1, 0).clone()
    eng.mask_corrupt(length, dcols, tasks, 1, 0)
    logits = torch.empty((B * S, eng.logit_width), device="cuda")
    eng.forward(length, None, False, 1, 0, logits_out=logits)
    row = torch.zeros(eng.metrics_width, device="cuda")
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
    eng.backward(length, None, False, 1, 0)
    g1 = eng.grads.clone()
    r1 = row.clone()
    assert torch.isfinite(logits).all() and torch.isfinite(g1).all()
    # (1) padded elements never influence the loss or any gradient: garble them and repeat
    valid = (torch.arange(S, device="cuda")[None, :] <= length[:, None])
    for f, key in enumerate(m.keys):
        if cols[key]["type"] == "numerical":
            eng.modified[f][~valid] = 3.25
    eng.forward(length, None, False, 1, 0)
    eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
    eng.backward(length, None, False, 1, 0)
    assert torch.allclose(row[:-1], r1[:-1], rtol=1e-5, atol=1e-6)
    assert H.rel_l2(eng.grads.cpu().numpy(), g1.cpu().numpy()) < 1e-3  # split-K atomics reorder fp32 sums
    # (2) documents are independent: reversing the batch order gives the same loss
    perm = torch.arange(B - 1, -1, -1, device="cuda")
    length2 = length[perm].contiguous()
    cols2 = [c[perm].contiguous() for c in dcols]
    tasks2 = tasks[perm].contiguous()
    eng.mask_corrupt(length, dcols, tasks, 1, 0)
    mod2 = [c[perm].contiguous() for c in eng.modified]
    masks2 = [c[perm].contiguous() for c in eng.masks]
    eng.forward(length2, mod2, False, 1, 0)
    eng.loss(length2, cols2, masks2, row, 1.0 / B, False)
    assert torch.allclose(row[:-1], r1[:-1], rtol=2e-4, atol=1e-5)
    # (3) a few Adam steps on one batch reduce its loss
    m.compile(optimizer=Adam(learning_rate=1e-3, clipnorm=1.0))
    first = m.metrics_from_row(m.train_step(batch))["loss"]
    for _ in range(5):
        last = m.metrics_from_row(m.train_step(batch))["loss"]
    assert np.isfinite(last) and last < first


def test_tf32_operand_rounding_of_the_product_path():
    """What the tcgen05 GEMM does to fp32 operands below TF32's 10-bit mantissa: the TMA unit converts TFLOAT32 maps with
    round-to-nearest, ties to even (measured on a B200; the MMA alone would truncate) -- the rule ``oracle.tf32_round`` /
    ``oracle.emulate_tf32`` restate, on which the stated TF32 tolerances rest.  The products below are exact in fp32 under any rule,
    so the results identify it."""
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
    assert rules == {("A", "nearest, ties to even"), ("B", "nearest, ties to even")}, rules
