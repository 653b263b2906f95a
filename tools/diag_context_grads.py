"""Diagnostic for the open TF32 item of --context canvas: per-variable gradient error of the engine (both GEMM paths) against the float64 oracle on a golden batch, for several
context modes, with the worst variable's error broken down by 32-row / 32-column groups.  Usage: python tools/diag_context_grads.py"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flex_dm_b200.mfp import MFP  # noqa: E402
from flex_dm_b200.spec import make_input_columns  # noqa: E402
from oracle import mfp_oracle as O  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "crello_ctx_canvas.npz"))
cols = make_input_columns("crello", max_length=50)
batch = OrderedDict((k[3:], g[k]) for k in g.files if k.startswith("in/"))
L, seed, step = 2, 25, 0
tasks_np = g["tasks"]


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


for context in ("canvas", "id", None):  # same 27-row batch with and without a context token: is the defect tied to canvas or to the shape?
    params = O.init_params(cols, L, 256, 11, torch.float64, bias_scale=0.05, context=context)
    o = O.OracleMFP(cols, num_blocks=L, masking_method="random", dropout=0.1, l2=None, context=context)
    o.params = OrderedDict((k, v.clone()) for k, v in params.items())  # step_from applies Adam to o.params in place: keep `params` pristine for the engine
    draws = O.PhiloxDraws(seed, step)
    tasks = torch.as_tensor(tasks_np)
    targets, mod, masks = O.preprocess_for_train(o.to_torch(batch), o.input_columns, tasks, draws, "set")
    B, S = batch["left"].shape[:2]
    ref = o.step_from(targets, mod, masks, tasks, o.dropout_masks(draws, B, S))["grads"]
    for impl in (1, 0):
        for rep in range(2 if impl == 0 else 1):
            m = MFP(cols, num_blocks=L, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=0, context=context)
            m._pad_context = False
            m.set_weights({k: v.numpy().astype(np.float32) for k, v in params.items()})
            eng = m.engine
            eng.set_gemm_impl(impl)
            staged = m.stage(batch)
            _, _, length, dcols = m._bind(staged)
            t = torch.as_tensor(tasks_np).cuda()
            m._set_context(t)
            eng.mask_corrupt(length, dcols, t, seed, step)
            eng.forward(length, None, True, seed, step)
            row = torch.zeros(eng.metrics_width, device="cuda")
            eng.loss(length, dcols, eng.masks, row, 1.0 / B, True)
            eng.backward(length, None, True, seed, step)
            torch.cuda.synchronize()
            got = eng.get_weights(eng.grads)
            errs = sorted(((rel(got[k], ref[k].numpy()), k) for k in got if np.linalg.norm(ref[k].numpy()) > 0), reverse=True)
            print("context=%s impl=%d rep=%d T=%d worst:" % (context, impl, rep, B * S), ["%.2e %s" % (e, k.replace("model/", "")) for e, k in errs[:4]])
            if impl == 0 and rep == 0:
                e, k = errs[0]
                d = got[k].astype(np.float64) - ref[k].numpy()
                if d.ndim == 2:
                    r32 = np.sqrt((d ** 2).reshape(d.shape[0] // 32 if d.shape[0] % 32 == 0 else 1, -1).sum(1)) if d.shape[0] % 32 == 0 else None
                    c32 = np.sqrt((d ** 2).T.reshape(d.shape[1] // 32 if d.shape[1] % 32 == 0 else 1, -1).sum(1)) if d.shape[1] % 32 == 0 else None
                    print("   %s shape %s |err| by 32-row groups: %s" % (k, d.shape, None if r32 is None else np.round(r32, 2).tolist()))
                    print("   |err| by 32-column groups: %s" % (None if c32 is None else np.round(c32, 2).tolist()))
                    bad = np.argwhere(np.abs(d) > 0.05 * np.abs(ref[k].numpy()).max())
                    print("   entries off by > 5%% of the largest entry: %d; first: %s" % (len(bad), bad[:6].tolist()))
