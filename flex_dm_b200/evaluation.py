"""Host-side mirror of ``eval.py::evaluate`` (eval.py:35-112): the per-task evaluation loop of the paper's tables.

    from flex_dm_b200.evaluation import evaluate
    scores = evaluate(model, dataset, input_columns, task_mode="attr", group=("attr", ["opacity", "color", "font_family"]), num_iter=1)

For every batch it builds the test-time masks of the task, calls ``model(example, training=False, demo_args={"masks": ..., "num_iter": n})``
and ``LossLayer(input_columns)((example, prediction, masks)[, False, sort_flag])``, and accumulates ``<key>_score_num / _den``:

* ``task_mode="elem"`` (eval.py:67-73, one-shot models): the document is repeated S times and copy i hides element i in every
  sequence field (``eye(S)`` masks); like the reference this wants ``batch_size = 1`` (eval.py:141-143);
* ``task_mode`` in the attribute groups (``type / pos / attr / img / txt``, data/spec.py:364-377): every valid element of the
  group's fields is hidden; for rico and ``pos`` the loss layer sorts targets and predictions first (eval.py:104-106);
* ``task_mode="random"`` calls ``random_masking`` with keywords it does not accept (eval.py:58-64 vs masking.py:227-231), i.e. it
  raises in the reference; it raises here too.

All arithmetic runs in the engine (forward passes, loss / score kernel); this file is orchestration only.
"""
from collections import OrderedDict, defaultdict
from itertools import islice
from typing import Dict, Iterable, Optional, Sequence, Tuple

import numpy as np
import torch

from .masking import get_initial_masks, get_seq_mask
from .metrics import LossLayer
from .spec import get_dataset_name


def build_masks(example: Dict, input_columns: Dict, task_mode: str, group_keys: Optional[Sequence[str]] = None) -> Tuple[Dict, Dict]:
    """The masks (and, for ``elem``, the S-fold repeated example) of one evaluation batch: eval.py:52-93."""
    to_t = lambda v: v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
    example = {k: to_t(v) for k, v in example.items()}
    B, S = example["left"].shape[:2]
    seq_mask = get_seq_mask(example["length"], S)
    masks = get_initial_masks(input_columns, seq_mask)
    if task_mode == "random":
        raise TypeError("random_masking() got an unexpected keyword argument 'replace_prob' (eval.py:58-64 calls a signature masking.py does not have)")
    if task_mode == "elem":
        eye = torch.eye(S, dtype=torch.bool)
        repeated = {}
        for key, column in input_columns.items():
            if key not in example:
                continue
            repeated[key] = torch.repeat_interleave(example[key], S, dim=0)  # tf.repeat(example[key], S, axis=0)
            if column.get("is_sequence"):
                masks[key] = eye.repeat(B, 1) if B > 1 else eye  # B = 1 in the reference; per-document eye(S) otherwise
            elif not column.get("demo_only", False):
                masks[key] = torch.ones((B * S,), dtype=torch.bool)
        return repeated, masks
    if not group_keys:
        raise ValueError("task_mode %r needs the attribute group's keys" % task_mode)
    for key in group_keys:
        masks[key] = seq_mask
    return example, masks


def evaluate(model, dataset: Iterable[Dict], input_columns: Dict, task_mode: str, group: Optional[Tuple[str, Sequence[str]]] = None, num_iter: int = 1,
             steps: Optional[int] = None) -> "OrderedDict[str, float]":
    """eval.py:35-112.  ``group`` = (task name, keys of the attribute group), as ``get_attribute_groups(...).items()`` yields them."""
    group_name, group_keys = group if group else (task_mode, None)
    sort_pos = get_dataset_name(input_columns.keys()) == "rico"
    loss_layer = LossLayer(input_columns, device=model.device)
    total = defaultdict(float)
    for example in (islice(dataset, steps) if steps is not None else dataset):
        if example["left"].shape[1] == 0:
            continue
        batch, masks = build_masks(example, input_columns, task_mode, group_keys)
        B = batch["left"].shape[0]
        demo_args = {"masks": masks, "num_iter": num_iter}
        if getattr(model, "context", None) == "id":  # eval.py:99-101
            task_id = model.task_names.index(group_name)
            demo_args["tasks"] = torch.full((B,), task_id, dtype=torch.int32)
        prediction = model(batch, training=False, demo_args=demo_args)
        if sort_pos and task_mode == "pos":
            (scores,) = loss_layer((batch, prediction, masks), False, torch.ones((B,), dtype=torch.bool))
        else:
            (scores,) = loss_layer((batch, prediction, masks))
        for k, v in scores.items():
            total[k] += float(v)
    ans = OrderedDict()
    for k in input_columns:
        num_key, den_key = "%s_score_num" % k, "%s_score_den" % k
        if num_key in total:
            ans[k] = total[num_key] / total[den_key] if total[den_key] != 0.0 else float("nan")  # numpy 0/0 in the reference
    return ans
