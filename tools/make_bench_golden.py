#!/usr/bin/env python
"""CPU tool: the float64 oracle's loss of bench.py's FIRST step (initial weights, Philox key (seed 0, step 0), rank 0's batch 0) for the
bench configurations -> tests/golden/bench_step0.json.  bench.py compares the engine's step-0 loss with these numbers, so the
throughput it prints is tied to a parity-checked computation at the full shape.  Runs without a GPU (the engine library is only
asked for its variable inventory, which mfp_create builds on the host).

    python tools/make_bench_golden.py [config ...]        # default: 1 2 3 4
"""
import json
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from flex_dm_b200 import engine as E  # noqa: E402
from flex_dm_b200.mfp import init_weights  # noqa: E402
from flex_dm_b200.spec import make_input_columns, make_synthetic_batch  # noqa: E402
from oracle import mfp_oracle as O  # noqa: E402


def host_engine(cols, num_blocks):
    """An Engine object with the variable inventory but no device buffers (the trick of tests/test_cabi.py)."""
    E.load_library()
    eng = E.Engine.__new__(E.Engine)
    orig, zeros, dev = torch.cuda.is_available, torch.zeros, torch.cuda.device

    class _NoDev:
        def __init__(self, *_):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    try:
        torch.cuda.is_available = lambda: True
        torch.zeros = lambda *a, **k: zeros(*a, **{**k, "device": "cpu"})
        torch.cuda.device = _NoDev
        eng.__init__(cols, num_blocks=num_blocks, device="cpu")
    finally:
        torch.cuda.is_available, torch.zeros, torch.cuda.device = orig, zeros, dev
    return eng


def step0(config):
    w = bench.CONFIGS[config]
    cols = bench.input_columns_for(w)
    eng = host_engine(cols, w["L"])
    weights = init_weights(eng, 0)
    batch = make_synthetic_batch(cols, w["B"], w["S"], seed=0, lengths="full")  # bench.py: seed = 1000 * rank + batch index
    o = O.OracleMFP(cols, num_blocks=w["L"], masking_method=w["method"], dropout=0.1, l2=1e-2, dtype=torch.float64)
    o.params = OrderedDict((k, torch.tensor(v, dtype=torch.float64)) for k, v in weights.items())
    draws = O.PhiloxDraws(0, 0)
    inputs = o.to_torch(batch)
    tasks = torch.from_numpy(draws.tasks(w["B"], o.allowed_tasks))
    targets, modified, masks = O.preprocess_for_train(inputs, o.input_columns, tasks, draws, "set")
    with torch.no_grad():
        total, data_loss, losses, scores, metrics, _ = o.loss_from(o.params, targets, modified, masks, tasks, o.dropout_masks(draws, w["B"], w["S"]))
    return {"loss": float(total), "data_loss": float(data_loss), "l2_loss": float(total) - float(data_loss),
            "workload": w["name"], "B": w["B"], "S": w["S"], "L": w["L"], "seed": 0, "step": 0, "batch_seed": 0}


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    path = os.path.join(ROOT, "tests", "golden", "bench_step0.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for c in [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]:
        t0 = time.time()
        out["cfg%d" % c] = step0(c)
        print("cfg%d: %s (%.0f s)" % (c, out["cfg%d" % c], time.time() - t0), flush=True)
        json.dump(out, open(path, "w"), indent=1, sort_keys=True)
