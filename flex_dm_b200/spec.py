"""Column schema of the MFP hot path: what ``DataSpec.make_input_columns`` hands to ``MFP``.

The reference builds ``input_columns`` from a YAML spec plus vocabulary files that are external
downloads (``src/mfp/mfp/data/spec.py:144-211``, ``data/crello-spec.yml``, ``data/rico-spec.yml``).
The TFRecord reader is out of scope, but the *schema it emits* sizes every layer
of the hot path, so it is restated here as fixtures with synthetic vocabulary sizes.

Also restated (they are pure-Python helpers the hot path imports from ``data/spec.py``):
``ATTRIBUTE_GROUPS`` (spec.py:364-377), ``get_dataset_name`` (:380-385),
``get_attribute_groups`` (:388-390), ``get_valid_input_columns`` (:393-403).
"""
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np

ATTRIBUTE_GROUPS = {
    "rico": {
        "type": ["type"],
        "pos": ["left", "top", "width", "height"],
        "attr": ["icon", "clickable", "text_button"],
    },
    "crello": {
        "type": ["type"],
        "pos": ["left", "top", "width", "height"],
        "attr": ["opacity", "color", "font_family"],
        "img": ["image_embedding"],
        "txt": ["text_embedding"],
    },
}

# Element-type vocabulary of crello; index 0 is the lookup mask token '' (crello-spec.yml:37-44).
# The real order comes from vocabulary.json (external); names from helpers/svg_crello.py:70-78.
CRELLO_TYPES = ["", "svgElement", "textElement", "imageElement", "coloredBackground", "maskElement"]


def get_dataset_name(keys) -> str:
    return "rico" if "clickable" in keys else "crello"


def get_attribute_groups(keys) -> Dict[str, List[str]]:
    return ATTRIBUTE_GROUPS[get_dataset_name(keys)]


def get_valid_input_columns(input_columns: Dict, use_canvas: bool = False) -> Dict:
    outputs = OrderedDict()
    for key, column in input_columns.items():
        if key == "length":
            continue
        if column.get("demo_only", False):
            continue
        if not column["is_sequence"] and not use_canvas:
            continue
        outputs[key] = column
    return outputs


def _cat(input_dim, shape=(1,), is_sequence=True, primary_label=None):
    return {
        "type": "categorical",
        "input_dim": int(input_dim),
        "shape": tuple(shape),
        "is_sequence": is_sequence,
        "primary_label": primary_label,
    }


def _num(shape, is_sequence=True):
    return {"type": "numerical", "shape": tuple(shape), "is_sequence": is_sequence, "primary_label": None}


def _cond(values):
    return {"key": "type", "mask": [t in values for t in CRELLO_TYPES]}


def crello_input_columns(max_length: int = 50, font_vocab: int = 35) -> Dict:
    """crello-spec.yml column order (it fixes the fusion-sum order and the head order)."""
    c = OrderedDict()
    c["id"] = {"demo_only": True, "shape": (1,), "is_sequence": False, "primary_label": None}
    c["length"] = _cat(max_length, is_sequence=False)
    c["group"] = _cat(7, is_sequence=False)
    c["format"] = _cat(68, is_sequence=False)
    c["canvas_width"] = _cat(42, is_sequence=False)
    c["canvas_height"] = _cat(47, is_sequence=False)
    c["category"] = _cat(24, is_sequence=False)
    c["type"] = _cat(len(CRELLO_TYPES), primary_label=0)
    for k in ("left", "top", "width", "height"):
        c[k] = _cat(64)
    c["opacity"] = _cat(8)
    c["color"] = _cat(16, shape=(3,))
    c["color"]["loss_condition"] = _cond(["textElement", "coloredBackground"])
    c["image_embedding"] = _num((512,))
    c["image_embedding"]["loss_condition"] = _cond(["svgElement", "imageElement", "maskElement"])
    c["text_embedding"] = _num((512,))
    c["text_embedding"]["loss_condition"] = _cond(["textElement"])
    c["font_family"] = _cat(font_vocab)
    c["font_family"]["loss_condition"] = _cond(["textElement"])
    c["uuid"] = {"demo_only": True, "shape": (1,), "is_sequence": True, "primary_label": None}
    return c


def rico_input_columns(max_length: int = 50, type_vocab: int = 27, icon_vocab: int = 59, text_button_vocab: int = 26) -> Dict:
    """rico-spec.yml column order."""
    c = OrderedDict()
    c["length"] = _cat(max_length, is_sequence=False)
    for k in ("left", "top", "width", "height"):
        c[k] = _cat(64)
    c["clickable"] = _cat(2)
    c["type"] = _cat(type_vocab, primary_label=0)
    c["icon"] = _cat(icon_vocab)
    c["text_button"] = _cat(text_button_vocab)
    return c


def make_input_columns(dataset_name: str, max_length: int = 50) -> Dict:
    if dataset_name == "crello":
        return crello_input_columns(max_length)
    if dataset_name == "rico":
        return rico_input_columns(max_length)
    raise ValueError("Unknown dataset: %s" % dataset_name)


def make_synthetic_batch(
    input_columns: Dict,
    batch_size: int,
    seq_len: int,
    seed: int = 0,
    lengths: str = "full",
    fixed_lengths: Optional[np.ndarray] = None,
) -> Dict[str, np.ndarray]:
    """A batch shaped like ``DataSpec.parse_fn`` output (spec.py:255-287), SURVEY.md section 8d.

    ``length`` is zero-based ``(B,1)``; categorical sequence columns ``(B,S,C)`` int32 (padded positions hold 0);
    numerical ``(B,S,512)`` float32 with L2-normalised rows, exact zeros where the type gate excludes the element
    and on padded positions.  ``lengths``: ``"full"`` every document has S elements; ``"ragged"`` n ~ U{1..S}
    with at least one document of S elements (parse_sequence_example pads to the batch max).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    B, S = batch_size, seq_len
    if fixed_lengths is not None:
        n = np.asarray(fixed_lengths, dtype=np.int64).reshape(B)
    elif lengths == "full":
        n = np.full((B,), S, dtype=np.int64)
    elif lengths == "ragged":
        n = rng.integers(1, S + 1, size=(B,))
        n[rng.integers(0, B)] = S
    else:
        raise ValueError(lengths)
    valid = np.arange(S)[None, :] < n[:, None]
    batch: Dict[str, np.ndarray] = {}
    for key, column in input_columns.items():
        if column.get("demo_only", False):
            continue
        if key == "length":
            batch[key] = (n - 1).astype(np.int32).reshape(B, 1)
        elif not column["is_sequence"]:
            batch[key] = rng.integers(0, column["input_dim"], size=(B, 1)).astype(np.int32)
        elif column["type"] == "categorical":
            C = column["shape"][-1]
            lo = 1 if (key == "type" and get_dataset_name(input_columns.keys()) == "crello") else 0
            x = rng.integers(lo, column["input_dim"], size=(B, S, C)).astype(np.int32)
            batch[key] = x * valid[:, :, None]
        else:
            C = column["shape"][-1]
            x = rng.standard_normal(size=(B, S, C)).astype(np.float32)
            x /= np.linalg.norm(x, axis=-1, keepdims=True)
            batch[key] = (x * valid[:, :, None]).astype(np.float32)
    # exact zeros where the type gate says "not applicable" 
    for key, column in input_columns.items():
        if column.get("type") == "numerical" and "loss_condition" in column:
            cond = column["loss_condition"]
            gate = np.asarray(cond["mask"], dtype=bool)[batch[cond["key"]][..., 0]]
            batch[key] = (batch[key] * gate[:, :, None]).astype(np.float32)
    return batch
