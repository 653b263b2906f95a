"""world_size-2 gloo tests (CPU) of the document-sharded data-parallel logic in flex_dm_b200/parallel.py: sharding,
the single gradient all-reduce with inv_batch = 1/B_global, and the metric-row reduction.  The per-rank compute is the
float64 oracle (test infrastructure), so what is checked is exactly the host-side algebra the GPU path relies on."""
import os
import socket
from collections import OrderedDict

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flex_dm_b200 import parallel
from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from oracle import mfp_oracle as O

B, S, L = 5, 12, 1  # odd batch: uneven shards (3 + 2 documents)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(grads):
    return torch.cat([g.reshape(-1) for g in grads.values()])


def _local_step(cols, params, batch, tasks, b_global):
    """Oracle gradients of (1/B_global) * sum over this shard's documents, plus the additive metric row."""
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    inputs = {k: torch.as_tensor(v) for k, v in batch.items()}
    targets, modified, masks = O.preprocess_for_train(inputs, icols, torch.as_tensor(tasks), O.PhiloxDraws(3, 0))
    p = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in params.items())
    out = O.model_forward(p, modified, icols, L)
    total, losses, scores, _ = O.loss_layer(targets, out, masks, cols)
    b_local = inputs["length"].shape[0]
    (total * b_local / b_global).backward()  # loss_layer takes the mean over ITS batch (metrics.py:277)
    grads = OrderedDict((k, v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items())
    keys = list(O.get_valid_input_columns(cols).keys())
    row = []
    for k in keys:
        row += [float(losses[k]) * b_local / b_global, float(scores[k + "_score_num"]), float(scores[k + "_score_den"])]
    row += [float(total) * b_local / b_global, 123.0]  # data loss, replicated L2 column
    return _flat(grads), torch.tensor([row], dtype=torch.float64)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cols = make_input_columns("crello")
        params = O.init_params(cols, num_blocks=L, seed=1 + rank, bias_scale=0.05)  # deliberately different per rank
        flat = _flat(params)
        parallel.broadcast_parameters(dist, flat, 0)
        off = 0
        for k, v in params.items():
            params[k] = flat[off:off + v.numel()].reshape(v.shape).clone()
            off += v.numel()
        batch = make_synthetic_batch(cols, B, S, seed=2, lengths="ragged")
        tasks = np.zeros((B,), dtype=np.int32)
        lo, hi = parallel.shard_bounds(B, rank, world)
        shard = parallel.shard_documents(batch, rank, world)
        # the Philox draws are keyed by document position, so each rank replays the global draws and slices them
        g_full, row_full = _local_step(cols, params, batch, tasks, B)
        icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
        inputs = {k: torch.as_tensor(v) for k, v in batch.items()}
        targets, modified, masks = O.preprocess_for_train(inputs, icols, torch.as_tensor(tasks), O.PhiloxDraws(3, 0))
        sl = lambda d: {k: (v[lo:hi] if torch.is_tensor(v) and v.shape[:1] == (B,) else v) for k, v in d.items()}
        # what a rank really does (MFP.enable_data_parallel -> mfp_set_doc_offset): it corrupts ITS documents with Philox counters formed
        # from their global indices -- and gets exactly the global batch's draws for them
        shard_inputs = {k: torch.as_tensor(v) for k, v in shard.items()}
        t_l, mod_l, masks_l = O.preprocess_for_train(shard_inputs, icols, torch.as_tensor(tasks[lo:hi]), O.PhiloxDraws(3, 0, doc_offset=lo))
        for k in icols:
            if icols[k]["is_sequence"]:
                assert torch.equal(mod_l[k], sl(modified)[k]) and torch.equal(masks_l[k], sl(masks)[k]), k
        assert np.array_equal(O.PhiloxDraws(3, 0, doc_offset=lo).tasks(hi - lo, [0, 1, 3]), O.PhiloxDraws(3, 0).tasks(B, [0, 1, 3])[lo:hi])
        p = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in params.items())
        out_l = O.model_forward(p, mod_l, icols, L)
        total, losses, scores, _ = O.loss_layer(t_l, out_l, masks_l, cols)
        (total * (hi - lo) / B).backward()
        g = _flat(OrderedDict((k, v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()))
        keys = list(O.get_valid_input_columns(cols).keys())
        row = []
        for k in keys:
            row += [float(losses[k]) * (hi - lo) / B, float(scores[k + "_score_num"]), float(scores[k + "_score_den"])]
        row = torch.tensor([row + [float(total) * (hi - lo) / B, 123.0]], dtype=torch.float64)
        # the train step reduces one slice per backward stage, asynchronously; the pieces must add up to the one-shot all-reduce
        g_once = parallel.all_reduce_gradients(dist, g.clone())
        n = g.numel()
        cuts = [0, n // 5, n // 2, n]
        works = [parallel.all_reduce_gradient_slice(dist, g, cuts[i], cuts[i + 1]) for i in range(3)]
        assert parallel.all_reduce_gradient_slice(dist, g, 7, 7) is None  # an empty stage has nothing to exchange
        for w in works:
            w.wait()
        assert torch.equal(g, g_once)
        row = parallel.reduce_metric_rows(dist, row)
        out[rank] = (shard["length"].shape[0], float((g - g_full).abs().max() / g_full.abs().max()), float((row - row_full).abs().max()),
                     float(row[0, -1]))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for b in (1, 5, 8, 256, 513):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(b, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def test_two_rank_gradient_and_metric_reduction_matches_single_process():
    world = 2
    port = _free_port()
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    assert out[0][0] + out[1][0] == B and out[0][0] == 3
    for rank in range(world):
        n_docs, grad_err, row_err, l2 = out[rank]
        assert grad_err < 1e-10, (rank, grad_err)   # sum of shard gradients == global-batch gradient
        assert row_err < 1e-9, (rank, row_err)      # additive metrics
        assert l2 == 123.0                          # the replicated L2 column is not summed
