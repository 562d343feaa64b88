"""projection + dropout + residual + LayerNorm at the bench row count: fused kernel (emo_gemm_ln_res) vs emo_gemm + emo_ln_res_fwd"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from emo_disentanger_b200 import ops
M, d = int(os.environ.get("M", 74 * 2048)), 512
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for K in (512, 2048):
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16); w = (torch.randn(d, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias, g, b = (torch.randn(d, device="cuda") for _ in range(3))
    res = torch.randn(M, d, device="cuda").to(torch.bfloat16)
    y, s, b1 = (torch.empty(M, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    mean, rstd = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
    fused = timeit(lambda: ops.linear_ln_res_fwd(x, w, bias, res, g, b, y, mean, rstd, sum_out=s, drop_p=0.1, seed=5))
    t_g = timeit(lambda: ops.linear_fwd(x, w, b1, bias=bias, drop_p=0.1, seed=5))
    t_l = timeit(lambda: ops.ln_res_fwd(b1, res, g, b, y, mean, rstd, sum_out=b1))
    print("K=%4d: fused %.1f us | gemm %.1f + ln_res_fwd %.1f = %.1f us" % (K, fused, t_g, t_l, t_g + t_l))
