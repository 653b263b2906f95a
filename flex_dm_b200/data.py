"""Host -> device input pipeline of the B200 path: the stand-in for ``DataSpec.make_dataset(...).prefetch``
(reference: data/spec.py:231-253, which ends every dataset with ``dataset.prefetch(AUTOTUNE)``).

``DevicePrefetcher`` walks an iterable of host batches (numpy arrays or torch tensors, ideally pinned) and keeps
``depth`` batches in flight: while the engine computes step *i* on the current stream, the columns of step *i+1*
are copied into a second set of device buffers on a dedicated copy stream.  CUDA events order the two streams in
both directions (copy finished -> compute may read; compute finished -> buffers may be overwritten), so nothing
synchronises with the host.
"""
from typing import Dict, Iterable, Iterator, List, Optional

import numpy as np
import torch

ROWS_SUFFIX = "/rows"  # batch key of a packed column's element -> row map


def pack_batch(batch: Dict, input_columns: Dict) -> Dict:
    """The packed form of a ``DataSpec.parse_fn`` batch (data/spec.py:255-287).  Every numerical sequence column ``key`` -- float
    ``[B, S, C]``, e.g. the two 512-float embedding columns of crello, 4096 of the 4136 bytes of an element -- is replaced by

        batch[key]           float32 ``[n_rows, C]``: the rows of the elements that carry the field, document by document
        batch[key + "/rows"] int32 ``[B, S]``: element -> row, -1 where there is none

    An element carries the field iff it is a valid position (``s <= length``) and its type passes the column's ``loss_condition``
    (data/crello-spec.yml:88-121).  Everywhere else ``filter_padding`` (masking.py:24-53) overwrites the value with <UNUSED> before the
    model sees it and ``LossLayer`` gates the element out (metrics.py:251-267), so the packed batch gives bit-identical results
    (``mfp_set_packed_rows``) while the host -> device copy and the corruption pass move only the rows in use.  Other columns pass
    through; numpy in, numpy out."""
    out = dict(batch)
    length = np.asarray(batch["length"]).reshape(-1)
    types = None
    for key, column in input_columns.items():
        if not column.get("is_sequence") or column.get("type") != "numerical" or column.get("demo_only", False) or key not in batch:
            continue
        x = np.asarray(batch[key])
        B, S = x.shape[:2]
        carry = np.arange(S)[None, :] <= length[:, None]
        cond = column.get("loss_condition")
        if cond:
            if types is None:
                types = np.asarray(batch[cond["key"]])[..., 0]
            carry = carry & np.asarray(cond["mask"], dtype=bool)[types]
        rows = np.full((B, S), -1, dtype=np.int32)
        n = int(carry.sum())
        rows[carry] = np.arange(n, dtype=np.int32)
        out[key] = np.ascontiguousarray(x[carry], dtype=np.float32).reshape(n, x.shape[-1])
        out[key + ROWS_SUFFIX] = rows
    return out


def unpack_column(packed: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """Dense ``[B, S, C]`` column of a packed one (zeros where an element has no row), on the tensors' device."""
    B, S = rows.shape
    dense = torch.zeros((B, S, packed.shape[-1]), dtype=packed.dtype, device=packed.device)
    have = rows >= 0
    dense[have] = packed[rows[have].long()]
    return dense


class DevicePrefetcher:
    def __init__(self, model, batches: Iterable[Dict], depth: int = 2):
        self.model = model
        self.device = model.device
        self.depth = max(2, int(depth))
        self._it: Iterator = iter(batches)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.depth
        self._buffers: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.depth
        self._ready = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: List[Optional[torch.cuda.Event]] = [None] * self.depth
        self._pending: List[int] = []
        self._issued = 0
        self._last: Optional[int] = None
        self.bytes_per_batch = 0
        self._issue()

    def _issue(self) -> bool:
        try:
            batch = next(self._it)
        except StopIteration:
            return False
        k = self._issued % self.depth
        want = self.model.host_columns(batch)
        with torch.cuda.stream(self._copy_stream):
            if self._free[k] is not None:
                self._copy_stream.wait_event(self._free[k])
            # packed columns ([n_rows, C], n_rows changes from batch to batch) live in capacity buffers of B * S rows; the slot hands out views
            packed = {key for key in want if key + ROWS_SUFFIX in want}
            buffers = self._buffers[k]
            if buffers is None or any((key not in buffers) or (key not in packed and buffers[key].shape != tuple(t.shape)) or
                                      (key in packed and buffers[key].shape[0] < want[key + ROWS_SUFFIX].numel()) for key, t in want.items()):
                buffers = {key: torch.empty((want[key + ROWS_SUFFIX].numel(), t.shape[-1]) if key in packed else tuple(t.shape), dtype=t.dtype, device=self.device)
                           for key, t in want.items()}
                self._buffers[k] = buffers
            slot = {}
            nbytes = 0
            for key, t in want.items():
                dst = buffers[key][: t.shape[0]] if key in packed else buffers[key]
                dst.copy_(t, non_blocking=True)
                slot[key] = dst
                nbytes += t.numel() * t.element_size()
            self._slots[k] = slot
            self.bytes_per_batch = nbytes
            self._ready[k].record(self._copy_stream)
        self._pending.append(k)
        self._issued += 1
        return True

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        cur = torch.cuda.current_stream(self.device)
        if self._last is not None:  # the step that consumed the previous slot is enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self._free[self._last] = ev
        if not self._pending:
            raise StopIteration
        k = self._pending.pop(0)
        cur.wait_event(self._ready[k])
        self._issue()
        self._last = k
        return self._slots[k]
