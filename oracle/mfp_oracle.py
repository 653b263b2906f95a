"""TEST INFRASTRUCTURE (oracle) -- not product code.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product path never does.

PARITY STATUS: pinned against the reference's own Python, NOT against TensorFlow itself.  The reference
(CyberAgentAILab/flex-dm @ f2bcc9f) has no tests, golden vectors or fixtures, and its arithmetic lives in
tensorflow-gpu / tensorflow_probability (unpinned in requirements.txt:1,8; README.md:10 says TF 2.8) which cannot be
installed here (no wheel, Python 3.12, no network).  tests/golden/make_golden.py therefore imports the reference's
unmodified Python (mfp.models.mfp.MFP and everything under it) on top of a torch-backed stand-in for the slice of
the TF/Keras API it touches (oracle/tf_standin/) and records its outputs; tests/test_golden_reference.py checks this
file against those vectors: masking path bit-exact, logits / losses / scores / gradients / Adam update to 1e-10.
What remains unpinned ("parity unpinned" in the strict sense): the semantics of the TF/Keras primitives themselves
(each one a named constant/function below so it can be flipped in one place) -- the stand-in
restates them from the same recall as this file -- and TensorFlow's RNG streams.  The hand-derived known-answer tests
in tests/test_oracle_known_answers.py cover those primitives independently.

Everything is plain PyTorch-CPU tensor code in the dtype of the parameters (float64 = parity oracle,
float32 = the timed "port" CPU baseline).  All ``file:line`` citations are relative to
/root/reference/src/mfp/mfp/.
"""
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import philox

# ---- masking.py:8-15
MASK_VALUE = 10.0
NULL_VALUE = 0.0
MASK_PROB = 0.15
REPLACE_PROB = 0.1
UNCHANGE_PROB = 0.1
CHANGE_PROB = 1.0 - UNCHANGE_PROB
THRESH = REPLACE_PROB / CHANGE_PROB

# ---- Keras / TF 2.8 semantics
LN_EPS = 1e-3  # A1: LayerNormalization() default epsilon
CE_EPS = 1e-7  # A6: keras.backend.epsilon()
ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-7  # A5
NUM_HEADS = 8  # transformer.py:43 (default num_heads, never overridden: transformer.py:262-269)
SORT_KEYS = ["type", "left", "top", "width", "height"]  # tensor_utils.py:11

ATTRIBUTE_GROUPS = {  # data/spec.py:364-377
    "rico": {"type": ["type"], "pos": ["left", "top", "width", "height"], "attr": ["icon", "clickable", "text_button"]},
    "crello": {
        "type": ["type"],
        "pos": ["left", "top", "width", "height"],
        "attr": ["opacity", "color", "font_family"],
        "img": ["image_embedding"],
        "txt": ["text_embedding"],
    },
}


def get_dataset_name(keys):  # data/spec.py:380-385
    return "rico" if "clickable" in keys else "crello"


def get_attribute_groups(keys):  # data/spec.py:388-390
    return ATTRIBUTE_GROUPS[get_dataset_name(keys)]


def get_valid_input_columns(input_columns, use_canvas=False):  # data/spec.py:393-403
    out = OrderedDict()
    for key, column in input_columns.items():
        if key == "length" or column.get("demo_only", False):
            continue
        if not column["is_sequence"] and not use_canvas:
            continue
        out[key] = column
    return out


def get_task_names(input_columns):  # masking.py:18-21
    return ["random", "elem"] + list(get_attribute_groups(input_columns.keys()).keys())


def task_probs(task_names, masking_method):  # mfp.py:34-43
    used = masking_method.split("_")
    probs = [1.0 if n in used else 0.0 for n in task_names]
    total = sum(probs)
    assert total > 0.0
    return [p / total for p in probs]


# ----------------------------------------------------------------------------------------------- mask.py
def get_seq_mask(length: torch.Tensor, maxlen: Optional[int] = None) -> torch.Tensor:
    """architecture/mask.py:21-33 -- ``sequence_mask(length + 1)`` (length is zero-based)."""
    n = length.reshape(-1).to(torch.int64) + 1
    S = int(n.max()) if maxlen is None else int(maxlen)
    return torch.arange(S)[None, :] < n[:, None]


# ----------------------------------------------------------------------------------------------- RNG draws
class PhiloxDraws:
    """The B200 path's RNG contract, restated.  ``field`` is the index of the column among the
    sequence columns (``get_valid_input_columns`` order); token = b*S+s."""

    def __init__(self, seed: int, step: int = 0, doc_offset: int = 0):
        # doc_offset: global index of the batch's first document (mfp_set_doc_offset): counters follow global document / element indices
        self.seed, self.step, self.doc0 = int(seed), int(step), int(doc_offset)

    def tasks(self, B: int, allowed: List[int]) -> np.ndarray:
        x0, _, _, _ = philox.philox4x32_10(np.arange(B) + self.doc0, philox.FIELD_TASK, 0, 0, self.seed, self.step)
        return np.asarray(allowed, dtype=np.int32)[philox.mulhi_range(x0, len(allowed))]

    def uniforms(self, field: int, B: int, S: int):
        x0, x1, x2, _ = philox.philox4x32_10(np.arange(B * S) + self.doc0 * S, field, philox.STREAM_RANDOM_U, 0, self.seed, self.step)
        return [philox.u01(x).reshape(B, S) for x in (x0, x1, x2)]

    def rand_cat(self, field: int, B: int, S: int, C: int, input_dim: int) -> np.ndarray:
        t = (np.arange(B * S) + self.doc0 * S)[:, None]
        c = np.arange(C)[None, :]
        x0, _, _, _ = philox.philox4x32_10(t, field, philox.STREAM_RANDOM_CAT + c, 0, self.seed, self.step)
        return philox.mulhi_range(x0, input_dim).reshape(B, S, C).astype(np.int32)

    def rand_num(self, field: int, B: int, S: int, C: int) -> np.ndarray:
        assert C % 4 == 0
        t = (np.arange(B * S) + self.doc0 * S)[:, None]
        q = np.arange(C // 4)[None, :]
        x0, x1, x2, x3 = philox.philox4x32_10(t, field, philox.STREAM_RANDOM_NUM + q, 0, self.seed, self.step)
        z0, z1 = philox.box_muller(x0, x1)
        z2, z3 = philox.box_muller(x2, x3)
        z = np.stack([z0, z1, z2, z3], axis=-1).reshape(B, S, C)
        return (z * np.float32(0.1)).astype(np.float32)  # stddev=0.1, masking.py:91

    def elem_u(self, B: int) -> np.ndarray:
        x0, _, _, _ = philox.philox4x32_10(np.arange(B) + self.doc0, philox.FIELD_ELEM, 0, 0, self.seed, self.step)
        return philox.u01(x0)

    def shuffle_perm(self, lengths: np.ndarray, S: int) -> np.ndarray:
        """(B,S) source position of every output position: valid positions ordered by their Philox key (ties by position), padding in place."""
        B = len(lengths)
        perm = np.tile(np.arange(S, dtype=np.int64), (B, 1))
        x0, _, _, _ = philox.philox4x32_10(np.arange(B * S) + self.doc0 * S, philox.FIELD_SHUFFLE, 0, 0, self.seed, self.step)
        keys = x0.reshape(B, S).astype(np.int64)
        for b in range(B):
            n = int(lengths[b])
            perm[b, :n] = np.lexsort((np.arange(n), keys[b, :n]))
        return perm

    def pos_dropout_keep(self, shape, rate: float) -> np.ndarray:
        return philox.dropout_keep(int(np.prod(shape)), philox.SITE_POS_DROPOUT, rate, self.seed, self.step, self.doc0 * int(np.prod(shape[1:]))).reshape(shape)

    def dropout_keep(self, block: int, branch: int, shape, rate: float) -> np.ndarray:
        n = int(np.prod(shape))
        return philox.dropout_keep(n, philox.SITE_DROPOUT + 2 * block + branch, rate, self.seed, self.step, self.doc0 * int(np.prod(shape[1:]))).reshape(shape)


# ----------------------------------------------------------------------------------------------- masking.py
def apply_token(x: torch.Tensor, column: Dict, mask: torch.Tensor, token_type: str, random_values=None):
    """masking.py:68-95.  ``mask`` (B,S) bool; x (B,S,C)."""
    m = mask[..., None]
    if column["type"] == "categorical":
        mi = m.to(x.dtype)
        data = {"masked": column["input_dim"], "unused": column["input_dim"] + 1, "random": random_values}[token_type]
        return x * (1 - mi) + data * mi
    mf = m.to(x.dtype)
    data = {"masked": MASK_VALUE, "unused": NULL_VALUE, "random": random_values}[token_type]
    return x * (1.0 - mf) + data * mf


def filter_padding(inputs, input_columns, mask):
    """masking.py:24-53 -- <UNUSED> on padded elements and on fields the element's type excludes."""
    out = {}
    unused = ~mask
    for key, column in input_columns.items():
        x = inputs[key]
        if column["is_sequence"]:
            if "loss_condition" in column:
                cond = column["loss_condition"]
                m = torch.zeros_like(mask)
                for i, flag in enumerate(cond["mask"]):
                    if not flag:
                        m = m | (inputs[cond["key"]] == i)[..., 0]
                m = m | unused
            else:
                m = unused
            out[key] = apply_token(x, column, m, "unused")
        else:
            out[key] = x
    return out


def get_initial_masks(input_columns, mask):  # masking.py:56-65
    masks = {}
    for key, column in input_columns.items():
        if not column["is_sequence"]:
            masks[key] = torch.ones(mask.shape[0], dtype=torch.bool)
        else:
            masks[key] = torch.zeros_like(mask)
    return masks


def _seq_field_index(input_columns):
    return {k: i for i, k in enumerate(get_valid_input_columns(input_columns).keys())}


def random_masking(inputs, input_columns, mask, draws: PhiloxDraws):
    """masking.py:227-269 -- BERT-style 15% / 80-10-10."""
    B, S = mask.shape
    fidx = _seq_field_index(input_columns)
    modified, masks = {}, {}
    for key, column in input_columns.items():
        if not column["is_sequence"]:
            modified[key] = inputs[key]
            masks[key] = torch.ones(inputs[key].shape, dtype=torch.bool)
            continue
        u1, u2, u3 = [torch.from_numpy(u) for u in draws.uniforms(fidx[key], B, S)]
        mfp_mask = mask & (u1 < np.float32(MASK_PROB))
        chg_mask = mfp_mask & (u2 < np.float32(CHANGE_PROB))
        C = column["shape"][-1]
        if column["type"] == "categorical":
            rnd = torch.from_numpy(draws.rand_cat(fidx[key], B, S, C, column["input_dim"])).to(inputs[key].dtype)
        else:
            rnd = torch.from_numpy(draws.rand_num(fidx[key], B, S, C)).to(inputs[key].dtype)
        x = apply_token(inputs[key], column, chg_mask & (u3 >= np.float32(THRESH)), "masked")
        x = apply_token(x, column, chg_mask & (u3 < np.float32(THRESH)), "random", rnd)
        modified[key] = x
        masks[key] = mfp_mask
    return modified, masks


def select_single_element(mask, u: np.ndarray):
    """masking.py:98-113 (select_last=False): arr = int32(U * n_valid); one-hot; all-False if n_valid == 0."""
    length = mask.sum(dim=1).to(torch.float32)
    arr = (torch.from_numpy(u).to(torch.float32) * length).to(torch.int32)
    new_mask = torch.nn.functional.one_hot(arr.to(torch.int64), mask.shape[1]).to(torch.bool)
    return new_mask & (length > 0.0)[:, None]


def elem_masking(inputs, input_columns, mask, draws: PhiloxDraws):  # masking.py:136-155
    masks = get_initial_masks(input_columns, mask)
    selected = select_single_element(mask, draws.elem_u(mask.shape[0]))
    modified = {}
    for key, column in input_columns.items():
        if not column["is_sequence"]:
            modified[key] = inputs[key]
        else:
            modified[key] = apply_token(inputs[key], column, selected, "masked")
            masks[key] = selected
    return modified, masks


def feat_masking(inputs, input_columns, mask, feat_group):  # masking.py:116-133
    modified = {k: v.clone() for k, v in inputs.items()}
    masks = get_initial_masks(input_columns, mask)
    for key in feat_group:
        modified[key] = apply_token(modified[key], input_columns[key], mask, "masked")
        masks[key] = mask
    return modified, masks


def shuffle_inputs(inputs, perm: np.ndarray):
    """tensor_utils.py:47-76 with the permutation given: every tensor whose second axis is S is gathered along it."""
    S = perm.shape[1]
    idx = torch.as_tensor(perm, dtype=torch.int64)
    out = {}
    for key, val in inputs.items():
        if val.dim() >= 2 and val.shape[1] == S:
            full = idx.reshape(idx.shape + (1,) * (val.dim() - 2)).expand(-1, -1, *val.shape[2:])
            out[key] = torch.gather(val, 1, full)
        else:
            out[key] = val
    return out


def preprocess_for_train(inputs, input_columns, tasks: torch.Tensor, draws: PhiloxDraws, input_dtype: str = "set"):
    """mfp.py:95-138 (is_autoreg=False): optional shuffle; all masking variants are computed, then selected per document."""
    if input_dtype == "shuffled_set":  # mfp.py:104-105
        S_ = inputs[next(iter(get_valid_input_columns(input_columns)))].shape[1]
        inputs = shuffle_inputs(inputs, draws.shuffle_perm(inputs["length"].reshape(-1).numpy() + 1, S_))
    elif input_dtype == "sorted_set":  # mfp.py:106-107
        inputs, _ = sort_inputs(inputs, input_columns)
    groups = get_attribute_groups(input_columns.keys())
    S = inputs[next(iter(get_valid_input_columns(input_columns)))].shape[1]
    seq_mask = get_seq_mask(inputs["length"], S)
    filtered = filter_padding(inputs, input_columns, seq_mask)
    data = []
    modified, masks = random_masking(filtered, input_columns, seq_mask, draws)
    data.append(elem_masking(filtered, input_columns, seq_mask, draws))
    for group in groups.values():
        data.append(feat_masking(filtered, input_columns, seq_mask, group))
    for key in list(modified.keys()):
        for i, (mod_tmp, masks_tmp) in enumerate(data):
            cond = tasks == (i + 1)
            if input_columns[key]["is_sequence"]:
                cond = cond[..., None]
            modified[key] = torch.where(cond[..., None], mod_tmp[key], modified[key])
            if input_columns[key]["is_sequence"]:
                masks[key] = torch.where(cond, masks_tmp[key], masks[key])
    modified["task"] = tasks[..., None]
    return inputs, modified, masks


def preprocess_for_test(inputs, input_columns, masks, tasks=None):
    """mfp.py:72-92."""
    S = inputs[next(iter(get_valid_input_columns(input_columns)))].shape[1]
    seq_mask = get_seq_mask(inputs["length"], S)
    filtered = filter_padding(inputs, input_columns, seq_mask)
    modified = {}
    for key, column in input_columns.items():
        if not column["is_sequence"]:
            modified[key] = filtered[key]
            continue
        modified[key] = apply_token(filtered[key], column, masks[key], "masked")
    if tasks is None:
        tasks = torch.zeros(inputs["length"].shape[0])
    modified["task"] = tasks[..., None]
    return modified


# ----------------------------------------------------------------------------------------------- parameters
def canvas_columns(input_columns, context):
    """The canvas-level columns the encoder embeds when ``"canvas" in context`` (encoder.py:34-37): the non-sequence entries of
    get_valid_input_columns(input_columns, use_canvas=True), i.e. everything but ``length`` and the demo-only columns."""
    if context not in ("canvas", "canvas_add"):
        return OrderedDict()
    icols = OrderedDict((k, c) for k, c in input_columns.items() if not c.get("demo_only", False))
    out = OrderedDict((k, c) for k, c in get_valid_input_columns(icols, True).items() if not c["is_sequence"])
    assert len(out) > 0  # encoder.py:205-206
    return out


def variable_specs(input_columns, num_blocks=4, latent_dim=256, input_dtype="set", context=None) -> "OrderedDict[str, Tuple[tuple, str, bool]]":
    """name -> (shape, init, l2-regularised).  SURVEY.md Appendix B; names follow the reference's attribute
    paths (mfp.py:249, model.py:20,45,52, encoder.py:74-92, transformer.py:54-57,161-173,263, decoder.py:39)."""
    D = latent_dim
    v = OrderedDict()
    cols = get_valid_input_columns(input_columns)
    canvas_cols = canvas_columns(input_columns, context)
    for key, c in canvas_cols.items():  # encoder.py:34-37,72-79: with use_canvas the canvas-level columns are embedded like any categorical column
        v["model/encoder/input_layer/%s/embeddings" % key] = ((c["input_dim"] + 2, D), "uniform", True)
    for key, c in cols.items():
        base = "model/encoder/input_layer/%s" % key
        if c["type"] == "categorical":
            v[base + "/embeddings"] = ((c["input_dim"] + 2, D), "uniform", True)  # encoder.py:74-79
        else:
            v[base + "_special/embeddings"] = ((2, D), "uniform", True)  # encoder.py:82-87
            v[base + "/kernel"] = ((c["shape"][-1], D), "glorot", True)  # encoder.py:88-92
            v[base + "/bias"] = ((D,), "zeros", True)
    if input_dtype != "set":  # PositionEmbedding(latent_dim, maxlen=length input_dim): Embedding(maxlen + 1, D) (encoder.py:48-55, transformer.py:17-21)
        v["model/encoder/input_layer/const/embeddings/embeddings"] = ((input_columns["length"]["input_dim"] + 1, D), "uniform", True)
    if context == "id":  # encoder.py:96-103
        v["model/encoder/input_layer/task/embeddings"] = ((len(get_task_names(input_columns)), D), "uniform", True)
    elif context == "length":  # encoder.py:104-110
        v["model/encoder/input_layer/length/embeddings"] = ((input_columns["length"]["input_dim"], D), "uniform", True)
    for i in range(num_blocks):
        b = "model/blocks/seq2seq/seq2seq_%d" % i
        for d in ("dense_query", "dense_key", "dense_value", "combine_heads"):  # transformer.py:54-57
            v["%s/attn/%s/kernel" % (b, d)] = ((D, D), "glorot", True)
            v["%s/attn/%s/bias" % (b, d)] = ((D,), "zeros", True)
        v[b + "/mlp/layer_with_weights-0/kernel"] = ((D, 2 * D), "glorot", True)  # transformer.py:161-171
        v[b + "/mlp/layer_with_weights-0/bias"] = ((2 * D,), "zeros", True)
        v[b + "/mlp/layer_with_weights-1/kernel"] = ((2 * D, D), "glorot", True)
        v[b + "/mlp/layer_with_weights-1/bias"] = ((D,), "zeros", True)
        for n in ("norm1", "norm2"):  # transformer.py:172-173; LN gamma/beta are not regularised
            v["%s/%s/gamma" % (b, n)] = ((D,), "ones", False)
            v["%s/%s/beta" % (b, n)] = ((D,), "zeros", False)
    if context == "canvas":  # decoder.py:25-43: use_canvas = (context == "canvas") adds a head per canvas column (never in the loss: metrics.py:226)
        for key, c in canvas_cols.items():
            v["model/decoder/decoders/%s/kernel" % key] = ((D, c["shape"][-1] * c["input_dim"]), "glorot", True)
            v["model/decoder/decoders/%s/bias" % key] = ((c["shape"][-1] * c["input_dim"],), "zeros", True)
    for key, c in cols.items():
        units = c["shape"][-1] * c["input_dim"] if c["type"] == "categorical" else c["shape"][-1]  # decoder.py:33-37
        v["model/decoder/decoders/%s/kernel" % key] = ((D, units), "glorot", True)
        v["model/decoder/decoders/%s/bias" % key] = ((units,), "zeros", True)
    return v


def init_params(input_columns, num_blocks=4, latent_dim=256, seed=0, dtype=torch.float64, bias_scale=0.0, input_dtype="set", context=None):
    """Keras default initialisers (Appendix A9): Dense glorot-uniform / zero bias, Embedding U(-0.05, 0.05), LN ones/zeros.
    ``bias_scale`` > 0 perturbs biases / LN parameters so that parity tests exercise them."""
    rng = np.random.Generator(np.random.PCG64(seed))
    params = OrderedDict()
    for name, (shape, init, _) in variable_specs(input_columns, num_blocks, latent_dim, input_dtype, context).items():
        if init == "uniform":
            w = rng.uniform(-0.05, 0.05, size=shape)
        elif init == "glorot":
            limit = math.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-limit, limit, size=shape)
        elif init == "ones":
            w = np.ones(shape) + bias_scale * rng.standard_normal(size=shape)
        else:
            w = bias_scale * rng.standard_normal(size=shape)
        params[name] = torch.tensor(w.astype(np.float32), dtype=dtype)
    return params


# ----------------------------------------------------------------------------------------------- network
def layer_norm(x, gamma, beta):
    """A1: biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * gamma + beta


# ---- TF32 emulation (test utility): what the B200 product path computes in --------------------------------------------------------
# The engine's GEMMs (every Dense forward / dgrad / wgrad, QK^T and PV) multiply operands rounded to TF32 (10 explicit mantissa bits,
# round to nearest, ties to even: what the TMA unit's TFLOAT32 conversion does, measured on a B200 by
# tests/test_gpu_parity.py::test_tf32_operand_rounding_of_the_product_path; ``tf32_truncate`` is what an unrounded fp32 container
# would get from the MMA) and accumulate in fp32.  ``emulate_tf32()`` makes
# this oracle do the same in float64 -- forward and backward products -- so that tests can state what TF32 *predicts* for a quantity
# and tell rounding from defects (tests/test_oracle_known_answers.py::test_tf32_emulation_bounds_the_stated_tolerances).
def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round to nearest, ties to even (the measured rule of the TFLOAT32 tensor maps; ``cvt.rna`` would round ties away)."""
    bits = x.detach().to(torch.float32).contiguous().view(torch.int32)
    bits = (bits + 0xFFF + ((bits >> 13) & 1)) & ~0x1FFF  # sign-magnitude: rounding the raw bits rounds the magnitude
    return bits.view(torch.float32).to(x.dtype)


def tf32_truncate(x: torch.Tensor) -> torch.Tensor:
    """Drop the 13 low mantissa bits (what a TF32 MMA does to an fp32 container that was not rounded first)."""
    bits = x.detach().to(torch.float32).contiguous().view(torch.int32) & ~0x1FFF
    return bits.view(torch.float32).to(x.dtype)


_tf32_rounding = tf32_round


class _Tf32MatMul(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return _tf32_rounding(a) @ _tf32_rounding(b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        gr, ar, br = _tf32_rounding(g), _tf32_rounding(a), _tf32_rounding(b)
        if b.dim() == 2 and a.dim() > 2:  # Dense over (B, S, K): the weight gradient contracts over all tokens
            return gr @ br.transpose(-1, -2), ar.reshape(-1, a.shape[-1]).transpose(0, 1) @ gr.reshape(-1, g.shape[-1])
        return gr @ br.transpose(-1, -2), ar.transpose(-1, -2) @ gr


_matmul = torch.matmul


class emulate_tf32:
    """``with emulate_tf32(): ...`` -- every matrix product of the model runs on TF32-rounded operands (forward and backward);
    ``rounding`` = ``tf32_round`` (default), ``tf32_truncate`` or any operand-rounding function."""

    def __init__(self, rounding=None):
        self._rounding = rounding or tf32_round

    def __enter__(self):
        global _matmul, _tf32_rounding
        self._saved = (_matmul, _tf32_rounding)
        _matmul, _tf32_rounding = _Tf32MatMul.apply, self._rounding
        return self

    def __exit__(self, *exc):
        global _matmul, _tf32_rounding
        _matmul, _tf32_rounding = self._saved


def dense(x, p, name):
    return _matmul(x, p[name + "/kernel"]) + p[name + "/bias"]  # A12


def encoder_forward(p, inputs, input_columns, pos_keep=None, pos_rate=0.0, context=None):
    """architecture/encoder.py:147-265 with fusion="add"; context None / "id" / "length"."""
    cols = get_valid_input_columns(input_columns)
    dtype = next(iter(p.values())).dtype
    S = inputs[next(iter(cols))].shape[1]
    seq_mask = get_seq_mask(inputs["length"], S)
    seq = 0.0
    for key, c in cols.items():
        base = "model/encoder/input_layer/%s" % key
        if c["type"] == "categorical":
            x = p[base + "/embeddings"][inputs[key].to(torch.int64)]  # (B,S,C,D)  encoder.py:157
            x = x.sum(dim=2)  # encoder.py:160
        else:
            xin = inputs[key].to(dtype)
            is_masked = (inputs[key] == MASK_VALUE).all(dim=2)  # encoder.py:165
            is_unused = (inputs[key] == NULL_VALUE).all(dim=2)  # encoder.py:166
            special = p[base + "_special/embeddings"]
            x = _matmul(xin, p[base + "/kernel"]) + p[base + "/bias"]  # encoder.py:173
            x = torch.where(is_masked[..., None], special[0], x)  # encoder.py:174
            x = torch.where(is_unused[..., None], special[1], x)  # encoder.py:175
        seq = seq + x  # encoder.py:194-197
    pos_name = "model/encoder/input_layer/const/embeddings/embeddings"
    canvas = 0.0
    for key in canvas_columns(input_columns, context):  # encoder.py:156-160,182-183,198-199: embed, sum over the sub-target axis, add up
        canvas = canvas + p["model/encoder/input_layer/%s/embeddings" % key][inputs[key].to(torch.int64)].sum(dim=1)
    if context == "canvas_add":  # encoder.py:228-230
        seq = seq + canvas[:, None, :]
    elif context is not None:  # encoder.py:231-249: a special token in front of the sequence, one more valid position per document
        if context != "canvas":
            ids = inputs["task"] if context == "id" else inputs["length"]  # :234-242
            ids = ids[:, 0] if ids.dim() == 2 else ids
            canvas = p["model/encoder/input_layer/%s/embeddings" % ("task" if context == "id" else "length")][ids.to(torch.int64)]
        seq = torch.cat([canvas[:, None, :], seq], dim=1)  # :247-248
        seq_mask = get_seq_mask(inputs["length"] + 1, S + 1)  # :249
    if pos_name in p:
        # use_pos_token (encoder.py:251-252): PositionEmbedding tiled over the batch, under its own Dropout -- added AFTER a context token
        # was put in front, i.e. the token takes position 0 and element s position s + 1 (transformer.py:24-29 ranges over shape[1])
        B, P = seq.shape[0], seq.shape[1]
        emb = p[pos_name][:P][None].expand(B, -1, -1)
        seq = seq + dropout(emb, pos_keep, pos_rate)
    return seq, seq_mask


def mhsa_forward(p, prefix, x, mask):
    """architecture/transformer.py:33-99."""
    B, S, D = x.shape
    H, dh = NUM_HEADS, D // NUM_HEADS

    def heads(t):  # transformer.py:78-80
        return t.reshape(B, S, H, dh).permute(0, 2, 1, 3)

    q = heads(dense(x, p, prefix + "/dense_query"))
    k = heads(dense(x, p, prefix + "/dense_key"))
    v = heads(dense(x, p, prefix + "/dense_value"))
    score = _matmul(q, k.transpose(-1, -2))
    score = score / math.sqrt(float(dh))  # transformer.py:62-63
    m = mask.to(x.dtype)[:, None, None, :]
    score = score + (-1e9) * (1.0 - m)  # transformer.py:73
    w = torch.softmax(score, dim=-1)
    out = _matmul(w, v).permute(0, 2, 1, 3).reshape(B, S, D)
    return dense(out, p, prefix + "/combine_heads")


def dropout(x, keep, rate):
    """A10: inverted dropout; ``keep`` None = inference / rate 0."""
    if keep is None:
        return x
    return x * keep.to(x.dtype) * (1.0 / (1.0 - rate))


_relu_hook = None  # test utility: callable(block index, pre-activations) -> activations; lets a test record the FFN pre-activations or
#                    force single ReLU gates (tests/test_oracle_known_answers.py::test_relu_gate_at_the_tf32_rounding_edge)


def ffn_relu(i, pre):
    return torch.relu(pre) if _relu_hook is None else _relu_hook(i, pre)


def blocks_forward(p, x, mask, num_blocks, drop=None, rate=0.0, block_type="deepsvg"):
    """architecture/transformer.py:208-229 (DeepSVGBlock, pre-LayerNorm) or :187-205 (TransformerBlock, post-LayerNorm: --block_type
    transformer) stacked by Blocks.__call__ :272-280; no final norm."""
    for i in range(num_blocks):
        b = "model/blocks/seq2seq/seq2seq_%d" % i
        if block_type == "transformer":
            y = mhsa_forward(p, b + "/attn", x, mask)
            y = dropout(y, None if drop is None else drop[(i, 0)], rate)
            x = layer_norm(x + y, p[b + "/norm1/gamma"], p[b + "/norm1/beta"])
            y = ffn_relu(i, dense(x, p, b + "/mlp/layer_with_weights-0"))
            y = dense(y, p, b + "/mlp/layer_with_weights-1")
            y = dropout(y, None if drop is None else drop[(i, 1)], rate)
            x = layer_norm(x + y, p[b + "/norm2/gamma"], p[b + "/norm2/beta"])
            continue
        y = layer_norm(x, p[b + "/norm1/gamma"], p[b + "/norm1/beta"])
        y = mhsa_forward(p, b + "/attn", y, mask)
        y = dropout(y, None if drop is None else drop[(i, 0)], rate)
        x = x + y
        y = layer_norm(x, p[b + "/norm2/gamma"], p[b + "/norm2/beta"])
        y = ffn_relu(i, dense(y, p, b + "/mlp/layer_with_weights-0"))  # transformer.py:163-166
        y = dense(y, p, b + "/mlp/layer_with_weights-1")
        y = dropout(y, None if drop is None else drop[(i, 1)], rate)
        x = x + y
    return x


def decoder_forward(p, h, input_columns):
    """architecture/decoder.py:72-111."""
    B, S, _ = h.shape
    out = OrderedDict()
    for key, c in get_valid_input_columns(input_columns).items():
        y = dense(h, p, "model/decoder/decoders/%s" % key)
        if c["type"] == "categorical":
            out[key] = y.reshape(B, S, c["shape"][-1], c["input_dim"])
        else:
            out[key] = y.reshape(B, S, c["shape"][-1])
    return out


def context_dropout_layout(drop, length):
    """Dropout keep-masks are drawn per row of the B200 engine's ``[B, S, D]`` activations, where the context token of document b
    sits in row ``n_b = length[b] + 1`` (the first padding row) and element s in row s.  The reference puts the token in front
    (position 0) and element s at position s + 1: gather the masks into that order.  The padding positions behind the token have no
    engine row of their own and no effect on any result; they reuse the rows at their own index."""
    if drop is None:
        return None
    n = (length.reshape(-1) + 1).to(torch.int64)
    out = {}
    for key, keep in drop.items():
        B, S, _ = keep.shape
        src = torch.arange(-1, S).repeat(B, 1)  # position p >= 1 reads row p - 1
        src[:, 0] = n.clamp(max=S - 1)  # the token's row
        src = src.clamp(min=0)
        out[key] = torch.gather(keep, 1, src[:, :, None].expand(-1, -1, keep.shape[2]))
    return out


def model_forward(p, modified_inputs, input_columns, num_blocks, drop=None, rate=0.0, return_hidden=False, block_type="deepsvg", context=None):
    """models/model.py:26-30."""
    if context in ("id", "length", "canvas"):
        drop = context_dropout_layout(drop, modified_inputs["length"])
    h0, mask = encoder_forward(p, modified_inputs, input_columns, None if drop is None else drop.get("pos"), rate, context)
    h = blocks_forward(p, h0, mask, num_blocks, drop, rate, block_type)
    if context in ("id", "length", "canvas"):  # decoder.py:74-78: the heads read the element positions only
        h = h[:, 1:]  # (with "canvas" the decoder also predicts the canvas columns from the token; nothing on this path reads them)
    out = decoder_forward(p, h, input_columns)
    if return_hidden:
        return out, h0, h
    return out


# ----------------------------------------------------------------------------------------------- tensor_utils.py
def sort_inputs(inputs, input_columns, from_logits=False):
    """tensor_utils.py:14-44 -- lexicographic element sort by (type,left,top,width,height); padded last."""
    CONST = 100
    data = {}
    for key, column in input_columns.items():
        if key not in inputs:
            continue
        v = inputs[key]
        if column.get("is_sequence") and column.get("type") == "categorical":
            if from_logits:
                v = v.argmax(dim=-1)
            v = v.to(torch.int64)
        data[key] = v
    S = data[SORT_KEYS[0]].shape[1]
    invalid = ~get_seq_mask(inputs["length"], S)
    priority = torch.zeros(data[SORT_KEYS[0]].shape[:2], dtype=torch.int64)
    for key in SORT_KEYS:
        priority = priority * CONST + data[key][..., 0]
    priority = priority + invalid.to(torch.int64) * (CONST ** len(SORT_KEYS))
    indices = torch.argsort(priority, dim=-1, stable=True)  # tf.argsort default is stable=False; ties are equal keys
    out = {}
    for key, val in inputs.items():
        if key in input_columns and input_columns[key].get("is_sequence", False):
            idx = indices.reshape(indices.shape + (1,) * (val.dim() - 2)).expand(-1, -1, *val.shape[2:])
            out[key] = torch.gather(val, 1, idx)
        else:
            out[key] = val
    return out, indices


# ----------------------------------------------------------------------------------------------- metrics.py
def categorical_metric(y_true, logits):
    """metrics.py:36-49 + A6: softmax -> clip -> log -> softmax-CE on the logs; score = (argmax == y)."""
    prob = torch.softmax(logits, dim=-1)
    arg = prob.argmax(dim=-1)
    z = torch.log(torch.clamp(prob, CE_EPS, 1.0 - CE_EPS))
    loss = -torch.gather(z, -1, y_true.to(torch.int64)[..., None])[..., 0] + torch.logsumexp(z, dim=-1)
    score = (y_true.to(torch.int64) == arg).to(logits.dtype)
    return loss, score


def continuous_metric(y_true, y_pred):
    """metrics.py:52-57 + A7/A8: MSE over the last axis; score = 0.5*cos + 0.5 (Keras l2-normalises with eps 1e-12)."""
    loss = ((y_true - y_pred) ** 2).mean(dim=-1)

    def l2n(t):
        return t * torch.rsqrt(torch.clamp((t * t).sum(dim=-1, keepdim=True), min=1e-12))

    cos = (l2n(y_true) * l2n(y_pred)).sum(dim=-1)
    return loss, 0.5 * cos + 0.5


def loss_layer(y_true, y_pred, mfp_masks, input_columns, sort_flag: Optional[torch.Tensor] = None):
    """metrics.py:173-299.  Returns (total_loss, losses, scores{num,den}, metrics)."""
    valid = get_valid_input_columns(input_columns)
    dtype = next(v for k, v in y_pred.items() if k in valid).dtype
    if sort_flag is not None:  # metrics.py:180-211
        y_true_sort, _ = sort_inputs(y_true, valid)
        y_pred_l = dict(y_pred)
        y_pred_l["length"] = y_true["length"]
        y_pred_sort, _ = sort_inputs(y_pred_l, valid, from_logits=True)
        yt, yp = {}, {}
        for key in y_true.keys():
            column = input_columns.get(key, {"demo_only": True})
            if column.get("demo_only", False):
                continue
            if column["is_sequence"]:
                flag = sort_flag[:, None, None]
                yt[key] = torch.where(flag, y_true_sort[key], y_true[key])
                if column["type"] == "categorical":
                    flag = flag[:, None]
                yp[key] = torch.where(flag, y_pred_sort[key], y_pred_l[key])
            else:
                yt[key] = y_true[key]
                if key in y_pred_l:
                    yp[key] = y_pred_l[key]
        y_true, y_pred = yt, yp
    S = y_true[next(iter(valid))].shape[1]
    seq_mask = get_seq_mask(y_true["length"], S)

    loss_total = 0.0
    score_total = 0.0
    losses, scores, metrics = OrderedDict(), OrderedDict(), OrderedDict()
    for key, column in input_columns.items():
        if column.get("demo_only", False) or not column["is_sequence"]:
            continue
        pred = y_pred[key][:, :S]
        if column["type"] == "categorical":
            assert int(y_true[key].max()) <= column["input_dim"] - 1 and int(y_true[key].min()) >= 0  # :236-237
            loss, score = categorical_metric(y_true[key], pred)
        else:
            loss, score = continuous_metric(y_true[key].to(dtype), pred)
            loss = loss[..., None] * float(column["shape"][-1])  # :246-247
            score = score[..., None]
        w = mfp_masks[key][..., None].to(dtype)  # :251
        loss = loss * w
        score = score * w
        den = torch.ones_like(loss) * w
        if "loss_condition" in column:  # :256-261
            cond = column["loss_condition"]
            gate = torch.tensor(cond["mask"], dtype=dtype)[y_true[cond["key"]].to(torch.int64)]
            loss, score, den = loss * gate, score * gate, den * gate
        sw = seq_mask[:, :, None].to(dtype)  # :263-267
        loss = (loss * sw).sum(dim=1).sum(dim=1)
        score = (score * sw).sum(dim=1).sum(dim=1)
        den = (den * sw).sum(dim=1).sum(dim=1)
        loss = loss.mean()  # :277 average batch
        score, den = score.sum(), den.sum()
        normalized = torch.where(den == 0.0, torch.ones_like(score), score / torch.where(den == 0.0, torch.ones_like(den), den))
        score_total = score_total + normalized
        metrics[key + "_score"] = normalized
        scores[key + "_score_num"] = score
        scores[key + "_score_den"] = den
        losses[key] = loss
    for key, loss in losses.items():
        metrics[key + "_loss"] = loss
        loss_total = loss_total + loss
    metrics["total_score"] = score_total / len(input_columns)  # :298
    return loss_total, losses, scores, metrics


def l2_regulariser(p, specs, l2):
    """architecture/utils.py:8-22 + A3: l2 * sum(w^2) over every Dense kernel+bias and Embedding table."""
    total = 0.0
    for name, (_, _, reg) in specs.items():
        if reg:
            total = total + l2 * (p[name] ** 2).sum()
    return total


def merge_inputs_and_prediction(inputs, input_columns, masks, prediction):
    """mfp.py:46-69."""
    out = dict(prediction)
    for key, column in input_columns.items():
        if column.get("demo_only", False):
            if key in inputs:
                out[key] = inputs[key]
            continue
        if not column["is_sequence"]:
            out[key] = inputs[key]
        elif key not in masks:
            continue
        elif column["type"] == "numerical":
            out[key] = torch.where(masks[key][..., None], prediction[key], inputs[key].to(prediction[key].dtype))
        else:
            gt = torch.nn.functional.one_hot(inputs[key].to(torch.int64), column["input_dim"]).to(prediction[key].dtype)
            out[key] = torch.where(masks[key][..., None, None], prediction[key], gt)
    return out


def iterative_decode(params, masks, inputs, input_columns, modified_inputs, num_iter, num_blocks):
    """mfp.py:141-207 -- MaskGIT-like decoding: ``num_iter`` forward passes; after each, the categorical predictions whose
    confidence (mean over sub-targets of the max softmax probability) reaches the document's top-k threshold are written
    back into the inputs and unmasked.  The reference compares a (B, S) confidence with a (B,) threshold (:184), which only
    broadcasts for B = 1; per-document thresholds (``threshold[:, None]``) are the same thing there and defined for any B."""
    masks = dict(masks)
    S = inputs[next(iter(get_valid_input_columns(input_columns)))].shape[1]
    seq_mask = get_seq_mask(inputs["length"], S)
    filtered = filter_padding(inputs, input_columns, seq_mask)
    cat_keys = [k for k, v in input_columns.items() if v["is_sequence"] and v.get("type", None) == "categorical"]
    num_masked = sum(masks[k].numpy().astype("int").sum(-1) for k in cat_keys)
    num_update = (num_masked / num_iter).round().astype("int")  # numpy: round half to even
    modified = dict(modified_inputs)
    final = None
    for i in range(num_iter):
        outputs = model_forward(params, modified, input_columns, num_blocks)
        if i == 0:
            final = OrderedDict(outputs)
        conf = {k: torch.where(masks[k], torch.softmax(outputs[k], dim=-1).max(dim=-1).values.mean(dim=-1), torch.zeros((), dtype=outputs[k].dtype))
                for k in cat_keys}
        conf_sorted = torch.sort(torch.cat([conf[k] for k in cat_keys], dim=-1), dim=-1, descending=True).values
        threshold = torch.stack([conf_sorted[b, k] for b, k in enumerate(num_update)])
        for key in cat_keys:
            pred = outputs[key].argmax(dim=-1).to(filtered[key].dtype)
            update = (conf[key] >= threshold[:, None]) & (conf[key] > 0)
            filtered[key] = torch.where(update[:, :, None], pred, filtered[key])
            masks[key] = torch.where(masks[key] == update, torch.zeros_like(masks[key]), masks[key])
            if i > 0:
                final[key] = torch.where(update[:, :, None, None], outputs[key], final[key])
        for key, column in input_columns.items():
            if column["is_sequence"]:
                modified[key] = apply_token(filtered[key], column, masks[key], "masked")
    for key, column in input_columns.items():  # the reference names image_embedding / text_embedding (:203-204)
        if column["is_sequence"] and column["type"] == "numerical":
            final[key] = outputs[key]
    return final


# ----------------------------------------------------------------------------------------------- optimiser
def clip_by_norm(g, clipnorm):
    """A4: tf.clip_by_norm per variable: g * c / max(||g||, c)."""
    n = torch.sqrt((g * g).sum())
    return g * (clipnorm / torch.maximum(n, torch.tensor(clipnorm, dtype=g.dtype)))


def adam_step(p, grads, m, v, t, lr=1e-4, clipnorm=1.0):
    """train.py:71-77 + A4/A5: per-variable clip, then TF Adam (eps outside the bias-corrected sqrt)."""
    alpha = lr * math.sqrt(1.0 - ADAM_BETA2**t) / (1.0 - ADAM_BETA1**t)
    for name in p:
        g = clip_by_norm(grads[name], clipnorm) if clipnorm is not None else grads[name]
        m[name] = ADAM_BETA1 * m[name] + (1.0 - ADAM_BETA1) * g
        v[name] = ADAM_BETA2 * v[name] + (1.0 - ADAM_BETA2) * g * g
        p[name] = p[name] - alpha * m[name] / (torch.sqrt(v[name]) + ADAM_EPS)


class OracleMFP:
    """The reference's MFP train/eval step (mfp.py:210-347 + Keras default train_step) on CPU."""

    def __init__(self, input_columns, num_blocks=4, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2,
                 seed=0, dtype=torch.float64, learning_rate=1e-4, clipnorm=1.0, bias_scale=0.0, block_type="deepsvg", input_dtype="set",
                 context=None):
        self.block_type = block_type
        self.context = context
        self.input_dtype = input_dtype
        self.input_columns = OrderedDict((k, v) for k, v in input_columns.items() if not v.get("demo_only", False))
        self.all_columns = input_columns
        self.num_blocks, self.latent_dim, self.rate, self.l2 = num_blocks, latent_dim, dropout, l2
        self.dtype = dtype
        self.specs = variable_specs(input_columns, num_blocks, latent_dim, input_dtype, context)
        self.params = init_params(input_columns, num_blocks, latent_dim, seed, dtype, bias_scale, input_dtype, context)
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.params.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.params.items())
        self.t = 0
        self.lr, self.clipnorm = learning_rate, clipnorm
        self.task_names = get_task_names(input_columns)
        probs = task_probs(self.task_names, masking_method)
        self.allowed_tasks = [i for i, pr in enumerate(probs) if pr > 0.0]
        self.sort_pos = get_dataset_name(input_columns.keys()) == "rico"  # mfp.py:293-296

    def to_torch(self, batch):
        return {k: torch.as_tensor(v) for k, v in batch.items()}

    def dropout_masks(self, draws: Optional[PhiloxDraws], B, S):
        if draws is None or self.rate == 0.0:
            return None
        keep = {(i, j): torch.from_numpy(draws.dropout_keep(i, j, (B, S, self.latent_dim), self.rate))
                for i in range(self.num_blocks) for j in (0, 1)}
        if self.input_dtype != "set":
            keep["pos"] = torch.from_numpy(draws.pos_dropout_keep((B, S, self.latent_dim), self.rate))
        return keep

    def loss_from(self, params, targets, modified, masks, tasks, drop):
        outputs = model_forward(params, modified, self.input_columns, self.num_blocks, drop, self.rate, block_type=self.block_type,
                                context=self.context)
        sort_flag = (tasks == self.task_names.index("pos")) if self.sort_pos else None  # mfp.py:335-340
        data_loss, losses, scores, metrics = loss_layer(targets, outputs, masks, self.all_columns, sort_flag)
        reg = l2_regulariser(params, self.specs, self.l2) if self.l2 is not None else 0.0
        return data_loss + reg, data_loss, losses, scores, metrics, outputs

    def train_step(self, batch, seed=0, step=0, training_dropout=True, doc_offset=0):
        """One Keras default train_step: sample tasks, corrupt, forward, loss (+L2), backward, clip, Adam."""
        inputs = self.to_torch(batch)
        draws = PhiloxDraws(seed, step, doc_offset)
        B = inputs["length"].shape[0]
        tasks = torch.from_numpy(draws.tasks(B, self.allowed_tasks))
        targets, modified, masks = preprocess_for_train(inputs, self.input_columns, tasks, draws, self.input_dtype)
        S = masks[next(iter(get_valid_input_columns(self.input_columns)))].shape[1]
        drop = self.dropout_masks(draws, B, S) if training_dropout else None
        return self.step_from(targets, modified, masks, tasks, drop)

    def step_from(self, targets, modified, masks, tasks, drop):
        params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in self.params.items())
        total, data_loss, losses, scores, metrics, outputs = self.loss_from(params, targets, modified, masks, tasks, drop)
        total.backward()
        grads = OrderedDict((k, v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in params.items())
        self.t += 1
        adam_step(self.params, grads, self.m, self.v, self.t, self.lr, self.clipnorm)
        total, data_loss = total.detach(), data_loss.detach()
        return {"loss": float(total), "data_loss": float(data_loss), "losses": {k: float(v.detach()) for k, v in losses.items()},
                "scores": {k: float(v.detach()) for k, v in scores.items()}, "metrics": {k: float(v.detach()) for k, v in metrics.items()},
                "grads": grads, "outputs": outputs}
