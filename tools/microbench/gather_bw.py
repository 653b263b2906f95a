"""mfp_gather_documents alone at the cfg2 batch shape (256 documents x 128 elements of the crello columns), CUDA-event timed:
algorithmic bytes = read + write of every column of the batch.  Usage: python tools/microbench/gather_bw.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flex_dm_b200 import engine  # noqa: E402

B, S, N = 256, 128, 1024
widths = [1, 1, 1, 1, 1, 1, 3, 512, 512, 1]  # type, left, top, width, height, opacity, color, image / text embedding, font_family
rng = np.random.default_rng(0)
doc_len = rng.integers(1, S + 1, N).astype(np.int32)
doc_len[: N // 2] = S
starts = np.zeros(N, dtype=np.int64)
starts[1:] = np.cumsum(doc_len)[:-1]
total = int(doc_len.sum())
dev = torch.device("cuda")
src = [torch.randn((total, w), device=dev) if w == 512 else torch.randint(0, 64, (total, w), dtype=torch.int32, device=dev) for w in widths]
d_start, d_len = torch.from_numpy(starts).to(dev), torch.from_numpy(doc_len).to(dev)
results = {}
for name, pick in (("full-length documents", np.arange(B)), ("ragged documents", np.arange(N // 2, N // 2 + B)), ("shuffled mix", rng.permutation(N)[:B])):
    idx = torch.from_numpy(pick.astype(np.int32)).to(dev)
    for _ in range(3):
        out = engine.gather_documents(src, [0] * len(src), d_start, d_len, idx, S)
    # check against torch indexing
    b = 5
    n = int(doc_len[pick[b]])
    for c, o in zip(src, out):
        assert torch.equal(o[b, :n], c[starts[pick[b]]:starts[pick[b]] + n]) and (o[b, n:] == 0).all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 50
    e0.record()
    for _ in range(reps):
        out = engine.gather_documents(src, [0] * len(src), d_start, d_len, idx, S)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    written = sum(o.numel() * 4 for o in out)
    read = int(doc_len[pick].sum()) * sum(widths) * 4
    results[name] = {"ms": ms, "bytes_written": written, "bytes_read": read, "GB/s": (written + read) / ms / 1e6}
print(json.dumps(results))
