#!/usr/bin/env python
"""N-GPU probe (torchrun): what the gradient all-reduce of the MFP step costs and which transport is cheapest.
Times, with CUDA events on the compute stream (max over ranks): NCCL all_reduce of the flat gradient buffer (11.25 MB fp32), the same
through torch symmetric memory (one-shot / two-shot / multimem NVLS kernels) when available, and the train step with and without the
staged-overlap path."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters, dev):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    n = 2812416  # crello flat gradient buffer (floats)
    g = torch.randn(n, device=dev)
    for _ in range(5):
        dist.all_reduce(g)
    out["nccl_all_reduce_ms"] = timed(lambda: dist.all_reduce(g), 50, dev)
    # back-to-back with a compute kernel in between (the step's pattern: the reduce sits between two compute kernels of the same stream)
    x = torch.randn(1 << 24, device=dev)

    def pattern():
        x.mul_(1.0001)
        dist.all_reduce(g)
        x.mul_(0.9999)

    def pattern_no():
        x.mul_(1.0001)
        x.mul_(0.9999)

    out["compute_allreduce_compute_ms"] = timed(pattern, 50, dev)
    out["compute_compute_ms"] = timed(pattern_no, 50, dev)
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        t.copy_(g)
        name = dist.group.WORLD.group_name
        for op in ("one_shot_all_reduce", "two_shot_all_reduce_", "multimem_all_reduce_", "multimem_one_shot_all_reduce"):
            try:
                fn = getattr(torch.ops.symm_mem, op)
                for _ in range(3):
                    fn(t, "sum", name)
                out["symm_mem_%s_ms" % op] = timed(lambda: fn(t, "sum", name), 50, dev)
            except Exception as e:  # noqa: BLE001
                out["symm_mem_%s_error" % op] = str(e)[:200]
        out["symm_mem_multicast"] = bool(getattr(hdl, "multicast_ptr", 0))
    except Exception as e:  # noqa: BLE001
        out["symm_mem_error"] = str(e)[:300]
    # the train step with the three gradient-exchange modes
    import bench

    for mode in ("none", "nccl", "nccl_overlap", "nvls"):
        bench.Workload.transport = "nccl" if mode != "nvls" else "nvls"
        wl = bench.Workload(2, world, rank, dev, dist)
        saved = None
        if mode == "none":
            import flex_dm_b200.mfp as M

            saved = M.all_reduce_gradients
            M.all_reduce_gradients = lambda d, gr: gr  # timing only: no exchange (results are wrong, rank-local)
        if mode == "nccl_overlap":
            wl.model._overlap = True
        for i in range(10):
            wl.step_resident(i)
        out["step_ms_" + mode] = timed(lambda: wl.step_resident(0), 100, dev)
        if saved is not None:
            M.all_reduce_gradients = saved
        del wl
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
