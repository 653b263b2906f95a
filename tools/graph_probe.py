#!/usr/bin/env python
"""Would a CUDA graph of the train step pay?  Captures ONE train step (fixed step counter: the replays reuse its RNG draws, so this is a
timing probe, not a training loop) with torch.cuda.graph and times replays against the normal launch path."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

docs = int(os.environ.get("DOCS", "0"))
bench.Workload.docs_override = docs
wl = bench.Workload(2, 1, 0, torch.device("cuda", 0), None)
for i in range(10):
    wl.step_resident(i)
torch.cuda.synchronize()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("docs per GPU %d: stream launches %.4f ms/step" % (wl.B, timed(wl.step_resident, 200)))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for i in range(3):
        wl.step_resident(0)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    wl.step_resident(0)
torch.cuda.synchronize()
print("docs per GPU %d: graph replays   %.4f ms/step" % (wl.B, timed(lambda i: g.replay(), 200)))
print("docs per GPU %d: stream launches %.4f ms/step (again)" % (wl.B, timed(wl.step_resident, 200)))
