#!/usr/bin/env python
"""Is the train step bound by the host's launch rate?  Pure CPU issue time of a short burst of steps (few enough launches to fit the
driver's launch queue, so the host never blocks on the GPU) beside the device time of the same burst."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = bench.Workload(int(os.environ.get("CFG", "2")), 1, 0, torch.device("cuda", 0), None)
for i in range(10):
    wl.step_resident(i)
torch.cuda.synchronize()
for burst in (2, 4, 8):
    cpu, tot = [], []
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(burst):
            wl.step_resident(i)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        cpu.append((t1 - t0) / burst * 1e3)
        tot.append((t2 - t0) / burst * 1e3)
    print("burst of %d steps: cpu issue %.3f ms/step (min %.3f), wall %.3f ms/step" % (burst, sum(cpu) / 5, min(cpu), sum(tot) / 5))
