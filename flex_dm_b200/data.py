"""Host -> device input pipeline of the B200 path: the stand-in for ``DataSpec.make_dataset(...).prefetch``
(reference: data/spec.py:231-253, which ends every dataset with ``dataset.prefetch(AUTOTUNE)``).

``DevicePrefetcher`` walks an iterable of host batches (numpy arrays or torch tensors, ideally pinned) and keeps
``depth`` batches in flight: while the engine computes step *i* on the current stream, the columns of step *i+1*
are copied into a second set of device buffers on a dedicated copy stream.  CUDA events order the two streams in
both directions (copy finished -> compute may read; compute finished -> buffers may be overwritten), so nothing
synchronises with the host.
"""
from typing import Dict, Iterable, Iterator, List, Optional

import torch


class DevicePrefetcher:
    def __init__(self, model, batches: Iterable[Dict], depth: int = 2):
        self.model = model
        self.device = model.device
        self.depth = max(2, int(depth))
        self._it: Iterator = iter(batches)
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Optional[Dict[str, torch.Tensor]]] = [None] * self.depth
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
            slot = self._slots[k]
            if slot is None or any(slot[key].shape != tuple(t.shape) for key, t in want.items()):
                slot = {key: torch.empty(tuple(t.shape), dtype=t.dtype, device=self.device) for key, t in want.items()}
                self._slots[k] = slot
            nbytes = 0
            for key, t in want.items():
                slot[key].copy_(t, non_blocking=True)
                nbytes += t.numel() * t.element_size()
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
