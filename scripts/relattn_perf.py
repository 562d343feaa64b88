"""Time the stage-1 relative-position attention at the stage-1 bench shape (B=64, T=512): tcgen05 path vs mma.sync."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops, _lib

B = int(os.environ.get("B", 64)); T = int(os.environ.get("T", 512)); H = 8; d = 512
dev = "cuda"
x = (torch.randn(B, T, 3 * d, device=dev) * 0.5).to(torch.bfloat16)
q, k, v = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
r = (torch.randn(T, H, 64, device=dev) * 0.5).to(torch.bfloat16)
rw, rr = 0.3 * torch.randn(H, 64, device=dev), 0.3 * torch.randn(H, 64, device=dev)
out = torch.empty(B, T, d, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
dout = torch.randn(B, T, d, device=dev).to(torch.bfloat16); dx = torch.empty_like(x)
dq, dk, dv = (dx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
dr = torch.zeros(T, H, 64, device=dev); drw = torch.zeros(H, 64, device=dev); drr = torch.zeros(H, 64, device=dev)
p = float(os.environ.get("P", 0.1))


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for tc in ((1, 0) if os.environ.get("AB", "1") == "1" else (1,)):
    _lib.lib().emo_attn_set_tc(tc)
    f = timeit(lambda: ops.relattn_fwd(q, k, v, r, rw, rr, out, lse, 0.125, p, 7))
    b = timeit(lambda: ops.relattn_bwd(q, k, v, r, rw, rr, out, dout, lse, dq, dk, dv, dr, drw, drr, 0.125, p, 7))
    print("%s B=%d T=%d p=%.1f: fwd %8.1f us | bwd %8.1f us" % ("tcgen05 " if tc else "mma.sync", B, T, p, f, b))
_lib.lib().emo_attn_set_tc(1)
