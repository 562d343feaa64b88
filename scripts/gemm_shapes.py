"""One launch of every GEMM of a Performer layer's train step (4 forward, 4 dgrad, 4 wgrad) at M = B*2048 tokens,
with the epilogues the model uses -- the target of the `ncu --set full` pass.  REPS>1 + no profiler: prints timings."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops

dev = "cuda"
M = int(os.environ.get("M", 74 * 2048)); REPS = int(os.environ.get("REPS", 1))
d, f = 512, 2048
bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(torch.bfloat16)
x, att, y1, hh = bf(M, d), bf(M, d), bf(M, d), torch.relu(bf(M, f))
wqkv, wo, w1, w2 = bf(3 * d, d), bf(d, d), bf(f, d), bf(d, f)
bq, bo, b1, b2 = (torch.randn(n, device=dev) for n in (3 * d, d, f, d))
qkv, s1, s2, h = bf(M, 3 * d), bf(M, d), bf(M, d), bf(M, f)
g2, g1, da, dqkv, dy1, datt, dx = bf(M, d), bf(M, d), bf(M, f), bf(M, 3 * d), bf(M, d), bf(M, d), bf(M, d)
dwqkv, dwo, dw1, dw2 = (torch.zeros(*w.shape, device=dev) for w in (wqkv, wo, w1, w2))
db1 = torch.zeros(f, device=dev)
cases = [
    ("fwd qkv       NT N=1536 K= 512 bias", 2.0 * M * 1536 * 512, lambda: ops.linear_fwd(x, wqkv, qkv, bias=bq)),
    ("fwd out-proj  NT N= 512 K= 512 bias+drop", 2.0 * M * 512 * 512, lambda: ops.linear_fwd(att, wo, s1, bias=bo, drop_p=0.1, seed=1)),
    ("fwd ffn1      NT N=2048 K= 512 bias+relu+drop", 2.0 * M * 2048 * 512, lambda: ops.linear_fwd(y1, w1, h, bias=b1, act=ops.ACT_RELU, drop_p=0.1, seed=2)),
    ("fwd ffn2      NT N= 512 K=2048 bias+drop", 2.0 * M * 2048 * 512, lambda: ops.linear_fwd(hh, w2, s2, bias=b2, drop_p=0.1, seed=3)),
    ("dgrad ffn2    NN N=2048 K= 512 relu-mask aux + colsum", 2.0 * M * 2048 * 512, lambda: ops.linear_dgrad(g2, w2, da, act=ops.ACT_RELU_MASK_BWD, aux=hh, ld_aux=f, aux_scale=1 / 0.9, colsum_out=db1)),
    ("dgrad ffn1    NN N= 512 K=2048 res", 2.0 * M * 2048 * 512, lambda: ops.linear_dgrad(da, w1, dy1, residual=g2, ld_res=d)),
    ("dgrad out     NN N= 512 K= 512", 2.0 * M * 512 * 512, lambda: ops.linear_dgrad(g1, wo, datt)),
    ("dgrad qkv     NN N= 512 K=1536 res", 2.0 * M * 1536 * 512, lambda: ops.linear_dgrad(dqkv, wqkv, dx, residual=g1, ld_res=d)),
    ("wgrad ffn2    TN 512x2048  K=M", 2.0 * M * 2048 * 512, lambda: ops.linear_wgrad(g2, hh, dw2)),
    ("wgrad ffn1    TN 2048x512  K=M", 2.0 * M * 2048 * 512, lambda: ops.linear_wgrad(da, y1, dw1)),
    ("wgrad out     TN 512x512   K=M", 2.0 * M * 512 * 512, lambda: ops.linear_wgrad(g1, att, dwo)),
    ("wgrad qkv     TN 1536x512  K=M", 2.0 * M * 1536 * 512, lambda: ops.linear_wgrad(dqkv, x, dwqkv)),
]
tot_f = tot_t = 0.0
for name, flops, fn in cases:
    if REPS == 1:
        fn()
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / REPS * 1e3
    tot_f += flops; tot_t += us
    print("%-52s %8.1f us %7.1f TFLOP/s" % (name, us, flops / us / 1e6), flush=True)
torch.cuda.synchronize()
if REPS > 1:
    print("layer total %.1f us  %.1f TFLOP/s" % (tot_t, tot_f / tot_t / 1e6))
