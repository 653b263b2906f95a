"""Host-side mirror of ``mfp/models/masking.py``: constants and the pure-Python helpers its callers import.

The tensor work of ``filter_padding`` / ``apply_token`` / ``random_masking`` / ``elem_masking`` /
``feat_masking`` (masking.py:24-155,227-269) runs inside one CUDA kernel (``csrc/masking.cu``), reached through
``Engine.mask_corrupt`` (train) and ``Engine.mask_for_test`` (demo / eval).
"""
from typing import Dict, List

import torch

from .spec import get_attribute_groups

# masking.py:8-15
MASK_VALUE = 10.0
NULL_VALUE = 0.0
MASK_PROB = 0.15
REPLACE_PROB = 0.1
UNCHANGE_PROB = 0.1
CHANGE_PROB = 1.0 - UNCHANGE_PROB
THRESH = REPLACE_PROB / CHANGE_PROB


def get_task_names(input_columns: Dict) -> List[str]:
    """masking.py:18-21."""
    task_names = ["random", "elem"]
    task_names += list(get_attribute_groups(input_columns.keys()).keys())
    return task_names


def get_seq_mask(length: torch.Tensor, maxlen: int = None) -> torch.Tensor:
    """architecture/mask.py:21-33 (``length`` is zero-based)."""
    n = length.reshape(-1).to(torch.int64) + 1
    S = int(n.max()) if maxlen is None else int(maxlen)
    return torch.arange(S, device=length.device)[None, :] < n[:, None]


def get_initial_masks(input_columns: Dict, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """masking.py:56-65: all-False masks for sequence columns, all-True placeholders for canvas columns."""
    masks = {}
    for key, column in input_columns.items():
        if column.get("demo_only", False):
            continue
        if not column["is_sequence"]:
            masks[key] = torch.ones(mask.shape[:1], dtype=torch.bool, device=mask.device)
        else:
            masks[key] = torch.zeros_like(mask, dtype=torch.bool)
    return masks
