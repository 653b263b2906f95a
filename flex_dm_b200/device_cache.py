"""``DataSpec.make_dataset(..., cache="device")``: the split parsed once and kept **ragged in HBM**, batches cut out of it on the GPU.

The reference's ``dataset.cache()`` (``data/spec.py:238-239``) keeps serialized records in host memory and re-parses them every epoch.
A B200 has room for the parsed crello / rico splits (a few GB of 180) many times over, so after the first pass a step's input costs one
HBM -> HBM gather (``mfp_gather_documents``, ``csrc/dataset.cu``) and ``B`` indices over PCIe; no host parsing, no 135 MB copy per step.

Batches are bit-identical to what ``make_dataset`` streams for the same ``seed`` (same index stream: shuffle / repeat / batch), which is what
``tests/test_gpu_input_pipeline.py`` asserts.  Demo-only byte-string columns are not cached.
"""
from collections import OrderedDict
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

from . import engine, io_lib


class DeviceCachedDataset:
    yields_device_batches = True  # MFP.fit feeds these to train_step directly (no host -> device prefetcher)

    def __init__(self, record_dataset, device: Optional[torch.device] = None, chunk: int = 2048):
        """``record_dataset``: the streaming ``RecordDataset`` of the same split (its shards, shuffle / repeat / batch settings and seed)."""
        self.source = record_dataset
        spec = record_dataset.spec
        self.spec = spec
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n = len(record_dataset)
        _, out_kind = spec._schema(False)
        self.seq_keys = [k for k, c in spec.columns.items() if c.get("is_sequence") and out_kind[k] != io_lib.OUT_SKIP]
        self.ctx_keys = [k for k, c in spec.columns.items() if not c.get("is_sequence") and out_kind[k] != io_lib.OUT_SKIP]
        if len(self.seq_keys) > engine.GATHER_MAX_COLUMNS:
            raise ValueError("too many sequence columns for the gather kernel")
        pieces: Dict[str, List[torch.Tensor]] = {k: [] for k in self.seq_keys}
        ctx: Dict[str, List[torch.Tensor]] = {k: [] for k in self.ctx_keys}
        lens: List[np.ndarray] = []
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            batch = spec.parse_records(record_dataset.pointers[lo:hi], record_dataset.lengths[lo:hi], pin_memory=True, strings=False)
            # elements of a document = its own step count (the `length` column may be looked-up, cut or absent: take it from the records)
            steps = spec.record_steps(record_dataset.pointers[lo:hi], record_dataset.lengths[lo:hi])
            lens.append(steps)
            S = batch[self.seq_keys[0]].shape[1] if self.seq_keys else 0
            keep = torch.from_numpy(np.arange(S)[None, :] < steps[:, None])
            for k in self.seq_keys:
                pieces[k].append(batch[k][keep].to(self.device, non_blocking=True))  # [elements of the chunk, C]
            for k in self.ctx_keys:
                ctx[k].append(batch[k].to(self.device, non_blocking=True))
        self.doc_len_host = np.concatenate(lens).astype(np.int32) if lens else np.zeros(0, np.int32)
        starts = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.doc_len_host, out=starts[1:])
        self.elements = OrderedDict()
        for k, v in pieces.items():
            width = int(np.prod(spec.columns[k].get("shape", (1,))))
            t = torch.cat(v).reshape(-1, width).contiguous() if v else None
            if t is None or t.numel() == 0:  # a split of empty documents: keep one dummy row so that the column has an address
                t = torch.zeros((1, width), dtype=torch.float32 if out_kind[k] == io_lib.OUT_FLOAT32 else torch.int32, device=self.device)
            self.elements[k] = t
        self.context = OrderedDict((k, torch.cat(v).contiguous()) for k, v in ctx.items())
        self.doc_start = torch.from_numpy(starts[:-1].copy()).to(self.device)
        self.doc_len = torch.from_numpy(self.doc_len_host).to(self.device)
        # what a padded step holds: the parse default put through the column's preprocessor (0 for every column of the two specs)
        self.pads = [spec.pad_word(k) for k in self.seq_keys]
        torch.cuda.synchronize(self.device)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in list(self.elements.values()) + list(self.context.values()))

    def __len__(self):
        return len(self.source)

    def batch(self, picked: List[int]) -> Dict[str, torch.Tensor]:
        idx_host = np.asarray(picked, dtype=np.int32)
        S = int(self.doc_len_host[idx_host].max()) if len(picked) else 0
        if self.source.pad_to is not None:
            if S > self.source.pad_to:
                raise ValueError("A document has more than pad_to=%d elements (%d)" % (self.source.pad_to, S))
            S = int(self.source.pad_to)
        idx = torch.from_numpy(idx_host).pin_memory().to(self.device, non_blocking=True)  # pinned staging: the copy does not block the host
        out = OrderedDict()
        long_idx = idx.long()
        for k, t in self.context.items():
            out[k] = t.index_select(0, long_idx)
        cols = engine.gather_documents(list(self.elements.values()), self.pads, self.doc_start, self.doc_len, idx, S)
        for k, t in zip(self.seq_keys, cols):
            out[k] = t
        return OrderedDict((k, out[k]) for k in self.spec.columns if k in out)

    def __iter__(self):
        from .dataspec import BatchIterator  # (dataspec imports this module lazily, too)

        return BatchIterator(self._generate())

    def _generate(self) -> Iterator[Dict[str, torch.Tensor]]:
        src = self.source
        rng = np.random.Generator(np.random.PCG64([src.seed, src._epoch]))
        src._epoch += 1
        picked: List[int] = []
        for i in src._index_stream(rng):
            picked.append(i)
            if len(picked) == src.batch_size:
                yield self.batch(picked)
                picked = []
        if picked:
            yield self.batch(picked)
