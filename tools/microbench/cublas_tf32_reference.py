import torch, time
torch.backends.cuda.matmul.allow_tf32 = True
dev = 'cuda'
def bench(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    # rotate inputs > L2? keep simple: same buffers (L2-warm) and also a flush variant
    flush = torch.empty(256*1024*1024//4, device=dev)
    ts = []
    for _ in range(n):
        flush.zero_()
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e)*1e3)
    ts.sort()
    return ts[len(ts)//2]
T = 32768
shapes = [("QKV fwd", (T,256),(256,768)), ("FFN1 fwd",(T,256),(256,512)), ("FFN2 fwd",(T,512),(512,256)), ("O fwd",(T,256),(256,256)), ("heads fwd",(T,256),(256,1408)),
          ("FFN1 dgrad",(T,512),(512,256)), ("QKV dgrad",(T,768),(768,256))]
for name, sa, sb in shapes:
    a = torch.randn(sa, device=dev); b = torch.randn(sb, device=dev); out = torch.empty((sa[0], sb[1]), device=dev)
    t = bench(lambda: torch.matmul(a, b, out=out))
    byt = 4*(a.numel()+b.numel()+out.numel())
    print("%-12s cuBLAS tf32 %.1f us  (%.0f GB/s algorithmic)" % (name, t, byt/t/1e3))
for name, k, m, n in [("FFN2 wgrad", T, 512, 256), ("QKV wgrad", T, 256, 768), ("heads wgrad", T, 256, 1408)]:
    x = torch.randn((k, m), device=dev); dy = torch.randn((k, n), device=dev); out = torch.empty((m, n), device=dev)
    t = bench(lambda: torch.matmul(x.t(), dy, out=out))
    byt = 4*(x.numel()+dy.numel()+out.numel())
    print("%-12s cuBLAS tf32 %.1f us  (%.0f GB/s algorithmic)" % (name, t, byt/t/1e3))
