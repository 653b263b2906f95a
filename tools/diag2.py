import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
from flex_dm_b200.mfp import MFP, Adam
from flex_dm_b200.spec import make_input_columns, make_synthetic_batch
from flex_dm_b200.data import DevicePrefetcher
variant = sys.argv[1]
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cols = make_input_columns("crello", max_length=128)
m = MFP(cols, num_blocks=4, masking_method="random", latent_dim=256, dropout=0.1, l2=1e-2, seed=0, device=dev)
m.compile(optimizer=Adam(1e-4, clipnorm=1.0))
host = [make_synthetic_batch(cols, 256, 128, seed=i, lengths="full") for i in range(4)]
need = [k for k, c in m.input_columns.items() if k == "length" or c["is_sequence"]]
pinned = [{k: torch.from_numpy(b[k]).pin_memory() for k in need} for b in host]
res = [m.stage(b) for b in pinned]
def timeit(fn, n=20, w=3):
    for i in range(w): fn(i)
    torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (t1 - t0) * 1e3 / n
if "sampler" in variant:
    sm = bench.ClockSampler(0); sm.start()
print("resident", timeit(lambda i: m.train_step(res[i % 4], staged=True), w=5))
if "sampler" in variant:
    print(sm.stop())
def gen():
    i = 0
    while True:
        yield pinned[i % 4]; i += 1
f = DevicePrefetcher(m, gen())
print("prefetcher", timeit(lambda i: m.train_step(next(f), staged=True)))
print("threads", threading.active_count())
