"""Which epilogue feature costs what on the N=512 shapes."""
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
for (N, K) in ((512, 512), (2048, 512)):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16)
    h = torch.relu(torch.randn(M, N, device=dev)).to(torch.bfloat16)
    cs = torch.zeros(N, device=dev)
    cases = [("plain", {}), ("bias", dict(bias=bias)), ("drop", dict(drop_p=0.1, seed=3)), ("res", dict(residual=res, ld_res=N)),
             ("bias+drop+res", dict(bias=bias, drop_p=0.1, seed=3, residual=res, ld_res=N)),
             ("relu-mask aux", dict(act=ops.ACT_RELU_MASK_BWD, aux=h, ld_aux=N, aux_scale=1.1)),
             ("relu-mask aux + colsum", dict(act=ops.ACT_RELU_MASK_BWD, aux=h, ld_aux=N, aux_scale=1.1, colsum_out=cs))]
    for name, kw in cases:
        for mode in (0, 1):
            _lib.lib().emo_gemm_debug(mode)
            us = timeit(lambda: ops.linear_fwd(a, w, out, **kw))
            print("N=%4d K=%4d %-24s %-12s %8.1f us  %7.1f TFLOP/s" % (N, K, name, "no-epilogue" if mode else "", us, 2.0 * M * N * K / us / 1e6), flush=True)
        _lib.lib().emo_gemm_debug(0)
