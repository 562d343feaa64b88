"""clock64 stamps of one worker thread; needs a library built with the stamps compiled in:
    EMO_NVCC_FLAGS=-DEMO_KERNEL_DBG_CLK python -m emo_disentanger_b200.build --force"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["EMO_FAVOR_DBG_CLK"] = str(buf.data_ptr())
from emo_disentanger_b200 import ops
B, T, H, d = 74, 2048, 8, 512
qkv = (torch.randn(B, T, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
q, k, v = (qkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
om = torch.randn(64, 64, device="cuda")
out = torch.empty(B, T, d, device="cuda", dtype=torch.bfloat16); den = torch.empty(B, T, H, device="cuda")
st = ops.favor_workspace(B, T, H, torch.bfloat16, "cuda")
for _ in range(3):
    ops.favor_fwd(q, k, v, om, out, den, seg_states=st)
torch.cuda.synchronize()
c = buf.cpu().tolist()
names = ["top", "batch1 done", "ssq done", "phi(k) done", "phi(q) done", "[B1]", "zsum done", "batch2 done", "P done", "[B2]", "batch3 done", "out done", "end"]
print([(names[i] if i < len(names) else i, c[i] - c[0]) for i in range(16) if c[i]])
