#!/usr/bin/env python
"""What HBM bandwidth can a SHORT kernel reach?  The step's kernels move 70-220 MB each (20-60 us); MEASURED_PEAKS.json's 6.5 TB/s is a
2 GiB copy.  Times torch copy_ (read + write bytes) per size, each launch on cold data (a rotating set of buffers larger than the 126 MB
L2), back to back like the step's kernels, and one launch at a time (ncu-like)."""
import json
import torch

dev = torch.device("cuda", 0)
out = []
for mb in (16, 33, 67, 100, 134, 268, 537, 1074, 2147):
    n = mb * 1000 * 1000 // 4
    k = max(2, int(600e6 // (n * 4)) + 1)  # rotate over > 126 MB per direction
    src = [torch.randn(n, device=dev) for _ in range(k)]
    dst = [torch.empty(n, device=dev) for _ in range(k)]
    for i in range(k):
        dst[i].copy_(src[i])
    torch.cuda.synchronize()
    iters = max(10, int(3e9 // (n * 8)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        dst[i % k].copy_(src[i % k])
    e1.record()
    torch.cuda.synchronize()
    back_to_back = 2 * n * 4 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
    single = []
    for i in range(10):
        torch.cuda.synchronize()
        e0.record()
        dst[i % k].copy_(src[i % k])
        e1.record()
        torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1))
    single.sort()
    out.append({"copy_MB_each_way": mb, "us_per_launch_back_to_back": 2 * n * 4 / back_to_back / 1e3, "GBps_back_to_back": back_to_back,
                "GBps_single_launch_median": 2 * n * 4 / (single[5] * 1e-3) / 1e9})
    print(out[-1], flush=True)
    del src, dst
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/copy_bw_by_size.json", "w"), indent=1)
