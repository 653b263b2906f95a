import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flex_dm_b200.mfp import MFP, Adam
from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from flex_dm_b200.data import DevicePrefetcher
cols = make_input_columns("crello", max_length=128)
m = MFP(cols, num_blocks=4, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=0)
m.compile(optimizer=Adam(1e-4, clipnorm=1.0))
host = [make_synthetic_batch(cols, 256, 128, seed=i, lengths="full") for i in range(4)]
need = [k for k, c in m.input_columns.items() if k == "length" or c["is_sequence"]]
pinned = [{k: torch.from_numpy(b[k]).pin_memory() for k in need} for b in host]
print("pinned?", all(t.is_pinned() for t in pinned[0].values()))
res = [m.stage(b) for b in pinned]
def timeit(fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (t1 - t0) * 1e3 / n
print("compute only   ms/step (gpu, host-enqueue):", timeit(lambda i: m.train_step(res[i % 4], staged=True)))
cs = torch.cuda.Stream()
slots = [{k: torch.empty_like(v, device="cuda") for k, v in pinned[0].items()} for _ in range(2)]
def copy_only(i):
    with torch.cuda.stream(cs):
        for k, v in pinned[i % 4].items(): slots[i % 2][k].copy_(v, non_blocking=True)
    torch.cuda.current_stream().wait_stream(cs)
print("copy only      ms/step:", timeit(copy_only))
def both_naive(i):
    with torch.cuda.stream(cs):
        for k, v in pinned[i % 4].items(): slots[i % 2][k].copy_(v, non_blocking=True)
    m.train_step(res[i % 4], staged=True)
def sync_both(i):
    both_naive(i)
    if i % 20 == 19: torch.cuda.current_stream().wait_stream(cs)
print("copy || compute (independent) ms/step:", timeit(sync_both))
def gen():
    i = 0
    while True:
        yield pinned[i % 4]; i += 1
f = DevicePrefetcher(m, gen())
print("prefetcher     ms/step:", timeit(lambda i: m.train_step(next(f), staged=True)))
rows_host = torch.empty((20, m.engine.metrics_width), dtype=torch.float32).pin_memory()
def with_d2h(i):
    row = m.train_step(next(f), staged=True)
    rows_host[i % 20].copy_(row, non_blocking=True)
print("prefetcher + D2H row ms/step:", timeit(with_d2h))
side = torch.cuda.Stream()
def with_d2h_side(i):
    row = m.train_step(next(f), staged=True)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        rows_host[i % 20].copy_(row, non_blocking=True)
print("prefetcher + D2H row on a side stream ms/step:", timeit(with_d2h_side))
