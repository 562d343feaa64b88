"""rows that share a sequence must share the result wherever they sit in the batch (dev aid)"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from helpers import rms_rel
import test_bench_config_gpu as TB
from emo_disentanger_b200 import ops
m, sd, om, tok, seg = TB._model(torch.bfloat16)
tok[37], seg[37] = tok[0], seg[0]
tok[73], seg[73] = tok[0], seg[0]
with torch.no_grad():
    h1, _ = m._forward_hidden(tok.cuda(), seg.cuda(), save=False)
    h2, _ = m._forward_hidden(tok.cuda(), seg.cuda(), save=False)
h1 = h1.view(TB.BB, TB.T, 512).float(); h2 = h2.view(TB.BB, TB.T, 512).float()
print("run-to-run", rms_rel(h1, h2), "row37 vs row0", rms_rel(h1[37], h1[0]), "row73 vs row0", rms_rel(h1[73], h1[0]))
# one FAVOR forward on identical rows
B, T, H = TB.BB, TB.T, TB.H
g = torch.Generator().manual_seed(3)
qkv = (torch.randn(1, T, 3 * 512, generator=g) * 0.7).to(torch.bfloat16).cuda().expand(B, T, 1536).contiguous()
q, k, v = (qkv[:, :, i * 512:(i + 1) * 512].unflatten(-1, (H, 64)) for i in range(3))
omega = torch.randn(64, 64, generator=g).cuda()
out = torch.empty(B, T, 512, device="cuda", dtype=torch.bfloat16)
den = torch.empty(B, T, H, device="cuda")
ws = ops.favor_workspace(B, T, H, torch.bfloat16, "cuda")
ops.favor_fwd(q, k, v, omega, out, den, seg_states=ws)
torch.cuda.synchronize()
bad = [(b, float((out[b].float() - out[0].float()).abs().max())) for b in range(B) if not torch.equal(out[b], out[0])]
print("favor rows differing from row 0:", bad[:10], len(bad))
# GEMM
x = torch.randn(1, T, 512, generator=g).to(torch.bfloat16).cuda().expand(B, T, 512).contiguous().view(B * T, 512)
w = (torch.randn(1536, 512, generator=g) * 0.05).to(torch.bfloat16).cuda()
y = torch.empty(B * T, 1536, device="cuda", dtype=torch.bfloat16)
ops.linear_fwd(x, w, y)
y = y.view(B, T, 1536)
print("gemm rows differing:", sum(0 if torch.equal(y[b], y[0]) else 1 for b in range(B)))
