import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from emo_disentanger_b200 import ops
B, T, V, d = 74, 2048, 329, 512
tok = torch.randint(0, V - 1, (B, T), device="cuda"); seg = torch.randint(0, 2, (B, T), device="cuda")
dout = torch.randn(B * T, d, device="cuda").to(torch.bfloat16)
det, des = torch.zeros(V, d, device="cuda"), torch.zeros(2, d, device="cuda")
for _ in range(3): ops.embed_bwd(tok, seg, dout, det, des, d ** 0.5, drop_p=0.1, seed=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.embed_bwd(tok, seg, dout, det, des, d ** 0.5, drop_p=0.1, seed=3)
e1.record(); torch.cuda.synchronize()
print("embed_bwd %.1f us" % (e0.elapsed_time(e1) / 10 * 1e3))
