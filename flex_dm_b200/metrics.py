"""Host-side mirror of ``mfp/models/metrics.py::LossLayer`` as a standalone callable (eval.py:44,104-108):

    loss_layer = LossLayer(input_columns)
    (scores,) = loss_layer((example, prediction, masks)[, training, sort_flag])

``scores`` holds ``<key>_score_num`` / ``<key>_score_den`` (metrics.py:287-288) as 0-d CPU tensors (``.numpy()``
works as in eval.py:109-110).  The arithmetic is ``csrc/loss.cu`` reached through ``Engine.loss``.
"""
from collections import OrderedDict
from typing import Dict, Optional

import torch

from .engine import Engine
from .spec import get_valid_input_columns


class LossLayer:
    def __init__(self, input_columns: Dict, name: str = "loss_layer", predict_context: bool = False, engine: Optional[Engine] = None, device=None):
        if predict_context:
            raise NotImplementedError("predict_context is never enabled by train.py / eval.py")
        self.name = name
        self._input_columns = input_columns
        self._valid_input_columns = get_valid_input_columns(input_columns)
        # the layer has no weights; a one-block engine carries the schema and the workspace
        self._engine = engine if engine is not None else Engine(input_columns, num_blocks=1, dropout=0.0, l2=None, device=device)
        self.keys = self._engine.keys
        self.losses: Dict[str, float] = {}
        self.metrics: Dict[str, float] = {}

    def _dev(self, x, dtype):
        x = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
        return x.to(self._engine.device).to(dtype).contiguous()

    def __call__(self, inputs, training: bool = False, sort_flag=None, ignore_sort: str = None):
        if ignore_sort is not None:
            raise NotImplementedError("ignore_sort is never used by train.py / eval.py")
        y_true, y_pred, mfp_masks = inputs
        eng = self._engine
        first = self._dev(y_true[self.keys[0]], torch.int32)
        B, S = int(first.shape[0]), int(first.shape[1])
        eng.bind(B, S)
        T = B * S
        length = self._dev(y_true["length"], torch.int32).reshape(-1)
        cols, masks = [], []
        logits = torch.zeros((T, eng.logit_width), dtype=torch.float32, device=eng.device)
        for f, key in enumerate(self.keys):
            c = self._valid_input_columns[key]
            cols.append(self._dev(y_true[key], torch.float32 if c["type"] == "numerical" else torch.int32))
            masks.append(self._dev(mfp_masks[key], torch.uint8))
            pred = self._dev(y_pred[key], torch.float32)[:, :S].reshape(T, -1)  # "Cut extra elements in prediction" (metrics.py:228)
            logits[:, eng.logit_offsets[f]:eng.logit_offsets[f] + pred.shape[1]] = pred
        flag = None
        if sort_flag is not None and torch.is_tensor(torch.as_tensor(sort_flag)):
            flag = self._dev(sort_flag, torch.uint8).reshape(-1)
        row = torch.zeros((eng.metrics_width,), dtype=torch.float32, device=eng.device)
        eng.loss(length, cols, masks, row, 1.0 / B, False, sort_flag=flag, logits_in=logits)
        r = row.cpu()
        scores = OrderedDict()
        total = 0.0
        self.losses, self.metrics = {}, {}
        for f, key in enumerate(self.keys):
            num, den = r[3 * f + 1], r[3 * f + 2]
            scores[key + "_score_num"] = num
            scores[key + "_score_den"] = den
            self.losses[key] = float(r[3 * f])
            norm = 1.0 if float(den) == 0.0 else float(num) / float(den)
            self.metrics[key + "_score"] = norm
            total += norm
        self.metrics["total_score"] = total / len(self._input_columns)
        self.metrics["loss"] = float(r[3 * len(self.keys)])
        return [scores]
