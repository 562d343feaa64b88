"""Time the FAVOR+ forward / backward kernels at the bench shape (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops

B = int(os.environ.get("B", 16)); T = 2048; H = 8; d = 512
dev = "cuda"
qkv = (torch.randn(B, T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = (qkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
om = torch.randn(64, 64, device=dev)
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16)
den = torch.empty(B, T, H, device=dev)
st = ops.favor_workspace(B, T, H, torch.bfloat16, dev)
print('nseg', st.shape[2])
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


f = timeit(lambda: ops.favor_fwd(q, k, v, om, out, den, seg_states=st))
b = timeit(lambda: ops.favor_bwd(q, k, v, om, out, dout, den, st, dq, dk, dv))
tok = B * T
print("favor fwd B=%d: %8.1f us  %7.1f GB/s algorithmic (4 KiB/token)" % (B, f, tok * 4096 / f / 1e3))
print("favor bwd B=%d: %8.1f us  %7.1f GB/s algorithmic (8 KiB/token)" % (B, b, tok * 8192 / b / 1e3))
