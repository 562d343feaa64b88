"""clock64 stamps of one worker thread; needs a library built with the stamps compiled in:
    EMO_NVCC_FLAGS=-DEMO_KERNEL_DBG_CLK python -m emo_disentanger_b200.build --force"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["EMO_ATTN_DBG_CLK"] = str(buf.data_ptr())
from emo_disentanger_b200 import ops
B, T, H, d = 16, 2048, 8, 512
qkv = (torch.randn(B, T, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
q, k, v = (qkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
out = torch.empty(B, T, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, H, T, device="cuda")
dout = torch.randn(B, T, d, device="cuda").to(torch.bfloat16); dqkv = torch.empty_like(qkv)
dq, dk, dv = (dqkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
p = float(os.environ.get("P", 0.0))
ops.attn_fwd(q, k, v, out, lse, 0.125, p, 7)
for _ in range(3):
    ops.attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, 0.125, p, 7)
torch.cuda.synchronize()
c = buf.cpu().tolist()
t0 = min(x for x in c if x > 0)
print("control:", [x - t0 if x else None for x in c[:16]])
print("worker0:", [x - t0 if x else None for x in c[16:]])
