"""LayerNorm forward / backward at the bench row count: achieved HBM GB/s (algorithmic bytes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200 import ops
R = int(os.environ.get("R", 74 * 2048)); dev = "cuda"
x = torch.randn(R, 512, device=dev).to(torch.bfloat16); dy = torch.randn(R, 512, device=dev).to(torch.bfloat16)
g, b = torch.randn(512, device=dev), torch.randn(512, device=dev)
y = torch.empty_like(x); mean = torch.empty(R, device=dev); rstd = torch.empty(R, device=dev)
dx = torch.empty_like(x); dxd = torch.empty_like(x); dg = torch.zeros(512, device=dev); db = torch.zeros(512, device=dev); cs = torch.zeros(512, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
f = timeit(lambda: ops.ln_fwd(x, g, b, y, mean, rstd))
w = timeit(lambda: ops.ln_bwd(dy, x, mean, rstd, g, dx, dg, db, dx_drop=dxd, drop_p=0.1, seed=3, dxsum=cs))
print("ln_fwd %8.1f us  %6.0f GB/s (2 KiB/row)" % (f, R * 2048 / f / 1e3))
print("ln_bwd %8.1f us  %6.0f GB/s (4 KiB/row: dy, x in; dx, dx_drop out)" % (w, R * 4096 / w / 1e3))
