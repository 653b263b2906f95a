"""Shared helpers of the parity tests: oracle <-> engine plumbing."""
from collections import OrderedDict

import numpy as np
import torch

from oracle import mfp_oracle as O

# Stated tolerances (BASELINE.json north_star: "within a stated fp32 tolerance"):
# the engine multiplies in TF32 (10-bit mantissa, operands rounded to nearest by the TMA unit) and accumulates in
# fp32; the oracle is float64.  Per-GEMM relative error is ~1e-3 of the operand norms.
LOGIT_ATOL = 2e-2      # max |logit - oracle| (logits are O(1)-O(10))
LOGIT_RTOL = 5e-3      # relative to the oracle's max |logit| of the field
LOSS_RTOL = 2e-3       # total / per-key loss
GRAD_REL_L2 = 5e-2     # ||g - g_oracle||_2 / ||g_oracle||_2 per variable (TF32 product path; tiny batches average least)
# With the fp32 SIMT GEMM (mfp_set_gemm_impl(1)) every other kernel is pinned at fp32 accuracy:
F32_LOGIT_ATOL = 2e-4
F32_LOSS_RTOL = 2e-5
F32_GRAD_REL_L2 = 2e-3  # fp32 accumulation order vs the float64 oracle (measured up to 6.2e-4)
WEIGHT_ATOL = 2e-6     # weights after one Adam step from identical gradients
# Compensated 3xTF32 on the tensor cores (mfp_set_gemm_impl(2): a_hi b_hi + a_lo b_hi + a_hi b_lo, each part 11 bits) carries ~21-22 mantissa
# bits per product against fp32's 24: same logit / loss bars as the fp32 path, gradients (long, cancelling sums over tokens) twice as wide
X3_GRAD_REL_L2 = 4e-3


def tolerances(impl):
    """(logit_atol, loss_rtol, grad_rel_l2) of a GEMM implementation: 0 = TF32 product path, 1 = fp32 SIMT, 2 = 3xTF32."""
    if impl == 0:
        return LOGIT_ATOL, LOSS_RTOL, GRAD_REL_L2
    return F32_LOGIT_ATOL, F32_LOSS_RTOL, (F32_GRAD_REL_L2 if impl == 1 else X3_GRAD_REL_L2)


def to_torch(batch):
    return {k: torch.as_tensor(v) for k, v in batch.items()}


def oracle_params_from_engine(engine, dtype=torch.float64):
    """Oracle parameter dict (reference variable names) from the engine's flat buffer."""
    return OrderedDict((k, torch.tensor(v, dtype=dtype)) for k, v in engine.get_weights().items())


def perturbed_weights(engine, seed=0, scale=0.05):
    """Keras-initialised weights with non-trivial biases / LayerNorm parameters so every term is exercised."""
    from flex_dm_b200.mfp import init_weights

    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    w = init_weights(engine, seed)
    for name in w:
        last = name.rsplit("/", 1)[-1]
        if last in ("bias", "beta"):
            w[name] = (scale * rng.standard_normal(w[name].shape)).astype(np.float32)
        elif last == "gamma":
            w[name] = (1.0 + scale * rng.standard_normal(w[name].shape)).astype(np.float32)
    return w


def oracle_train_inputs(cols, batch, tasks, seed, step):
    inputs = to_torch(batch)
    icols = OrderedDict((k, v) for k, v in cols.items() if not v.get("demo_only", False))
    return O.preprocess_for_train(inputs, icols, torch.as_tensor(tasks), O.PhiloxDraws(seed, step))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
