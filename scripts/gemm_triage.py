"""Perf triage of the tcgen05 GEMM: which pipeline stage bounds a shape (results are garbage in debug modes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops, _lib
dev = "cuda"; M = 32768
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (N, K) in ((2048, 512),):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for mode, name in ((0, "full"), (1, "no epilogue"), (2, "no MMA"), (3, "TMA loads only"), (6, "noMMA noTMEMld"), (10, "noMMA noStore"), (14, "noMMA epi math+smem only"), (14+32, "... and no math/smem"), (14+64, "... math, no sts")):
        _lib.lib().emo_gemm_debug(mode)
        us = timeit(lambda: ops.linear_fwd(a, w, out))
        print("N=%4d K=%4d %-16s %8.1f us  (%6.1f TFLOP/s-equivalent)" % (N, K, name, us, 2.0 * M * N * K / us / 1e6), flush=True)
    _lib.lib().emo_gemm_debug(0)
