"""Time the tcgen05 GEMM on the shapes of the stage-2 train step (CUDA events, L2-cold inputs rotate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops

dev = "cuda"
M = int(os.environ.get("M", 32768))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def run(name, N, K, out_dtype=torch.bfloat16, **kw):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    us = timeit(lambda: ops.linear_fwd(a, w, out, **kw))
    print("%-34s M=%d N=%4d K=%4d  %8.1f us  %7.1f TFLOP/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6), flush=True)


only = sys.argv[1] if len(sys.argv) > 1 else None
bias = torch.randn(2048, device=dev)
res = torch.randn(M, 2048, device=dev).to(torch.bfloat16)
cases = [
    ("plain bf16", 2048, 512, {}),
    ("plain fp32 out", 2048, 512, {"out_dtype": torch.float32}),
    ("bias", 2048, 512, {"bias": bias}),
    ("bias+relu", 2048, 512, {"bias": bias, "act": ops.ACT_RELU}),
    ("bias+relu+drop (ffn1)", 2048, 512, {"bias": bias, "act": ops.ACT_RELU, "drop_p": 0.1, "seed": 3}),
    ("qkv bias", 1536, 512, {"bias": bias[:1536]}),
    ("out-proj bias+drop+res", 512, 512, {"bias": bias[:512], "drop_p": 0.1, "seed": 3, "residual": res[:, :512], "ld_res": 2048}),
    ("ffn2 bias+drop+res", 512, 2048, {"bias": bias[:512], "drop_p": 0.1, "seed": 3, "residual": res[:, :512], "ld_res": 2048}),
]
for name, N, K, kw in cases:
    if only and only not in name:
        continue
    run(name, N, K, **kw)
# dgrad / wgrad
if not only:
    dy = torch.randn(M, 2048, device=dev).to(torch.bfloat16)
    w = (torch.randn(2048, 512, device=dev) * 0.05).to(torch.bfloat16)
    dx = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: ops.linear_dgrad(dy, w, dx))
    print("%-34s %8.1f us  %7.1f TFLOP/s" % ("dgrad ffn1 (NN) N=512 K=2048", us, 2.0 * M * 512 * 2048 / us / 1e6))
    x = torch.randn(M, 512, device=dev).to(torch.bfloat16)
    dw = torch.zeros(2048, 512, device=dev)
    us = timeit(lambda: ops.linear_wgrad(dy, x, dw))
    print("%-34s %8.1f us  %7.1f TFLOP/s" % ("wgrad ffn1 (TN) 2048x512 K=M", us, 2.0 * M * 512 * 2048 / us / 1e6))
