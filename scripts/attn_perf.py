"""Time the causal softmax attention kernels at the GPT-2 bench shape (CUDA events): tcgen05 (default) vs mma.sync."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops, _lib

B = int(os.environ.get("B", 16)); T = int(os.environ.get("T", 2048)); H = 8; d = 512
dev = "cuda"
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = (qkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
dout = torch.randn(B, T, d, device=dev).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
dq, dk, dv = (dqkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))


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


flop_f = 4.0 * 64 * B * H * T * (T + 1) / 2          # QK^T + PV over the causal half
for tc in ((1, 0) if os.environ.get("AB", "1") == "1" else (1,)):
    _lib.lib().emo_attn_set_tc(tc)
    for p in (0.0, 0.1):
        f = timeit(lambda: ops.attn_fwd(q, k, v, out, lse, 0.125, p, 7))
        b = timeit(lambda: ops.attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, 0.125, p, 7))
        print("%s p=%.1f B=%d T=%d: fwd %8.1f us %7.1f TFLOP/s | bwd %8.1f us %7.1f TFLOP/s" %
              ("tcgen05 " if tc else "mma.sync", p, B, T, f, flop_f / f / 1e6, b, 2.5 * flop_f / b / 1e6))
_lib.lib().emo_attn_set_tc(1)
