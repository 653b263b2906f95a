"""Time the engine's GEMM (mfp_debug_gemm) per operand layout, single CTAs vs CTA pairs (run twice: FLEXDM_GEMM_PAIR=0 / 1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from flex_dm_b200.engine import debug_gemm  # noqa: E402


def bench(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%-44s %8.2f us" % (name, e0.elapsed_time(e1) * 1000 / reps))


T = 32768
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
print("FLEXDM_GEMM_PAIR =", os.environ.get("FLEXDM_GEMM_PAIR", "(default on)"))
for N, K in [(256, 256), (768, 256), (256, 768), (512, 256), (256, 512)]:
    A, Wk, Wn = r(T, K), r(N, K), r(K, N)
    out = torch.empty(T, N, device="cuda")
    bench("fwd   A[T,%d] K-major x B[%d][%d] MN-major" % (K, K, N), lambda: debug_gemm(A, 0, Wn, 1, T, N, K, out=out))
    bench("dgrad A[T,%d] K-major x B[%d][%d] K-major" % (K, N, K), lambda: debug_gemm(A, 0, Wk, 0, T, N, K, out=out))
for M, N, s in [(256, 256, 74), (256, 768, 24), (512, 256, 37)]:
    X, dY = r(T, M), r(T, N)
    out = torch.zeros(M, N, device="cuda")
    bench("wgrad X[T,%d]^T x dY[T,%d] MN-major both, splits %d" % (M, N, s), lambda: debug_gemm(X, 1, dY, 1, M, N, T, splits=s, out=out))
