import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from emo_disentanger_b200 import ops
B, T, H = 74, 2048, 8
g = torch.Generator().manual_seed(3)
qkv = (torch.randn(B, T, 3 * 512, generator=g) * 0.7).to(torch.bfloat16).cuda()
q, k, v = (qkv[:, :, i * 512:(i + 1) * 512].unflatten(-1, (H, 64)) for i in range(3))
omega = torch.randn(64, 64, generator=g).cuda()
ref = None
ws = ops.favor_workspace(B, T, H, torch.bfloat16, "cuda")
nbad = 0
for it in range(60):
    out = torch.empty(B, T, 512, device="cuda", dtype=torch.bfloat16)
    den = torch.empty(B, T, H, device="cuda")
    ops.favor_fwd(q, k, v, omega, out, den, seg_states=ws)
    torch.cuda.synchronize()
    if ref is None:
        ref, refden = out, den
        continue
    d = (out.float() - ref.float()).view(B, T, H, 64)
    dd = (den - refden)
    if (d != 0).any() or (dd != 0).any():
        nbad += 1
        idx = (d != 0).any(-1).nonzero()       # (b, t, h)
        groups = {}
        for b, t, h in idx.tolist():
            groups.setdefault((b, h, t // 128), []).append(t % 128)
        for key, rows in list(groups.items())[:6]:
            b, h, c = key
            t0 = c * 128 + rows[0]
            ncols = int((d[b, t0, h] != 0).sum())
            print("it", it, "b,h,chunk", key, "rows", rows[0], "..", rows[-1], "n", len(rows), "cols differing in first row", ncols,
                  "den diff rows", int((dd[b, c * 128:(c + 1) * 128, h] != 0).sum()), "max", float(d[b, c*128:(c+1)*128, h].abs().max()),
                  "den ratio", [round(float(den[b, c * 128 + r, h] / refden[b, c * 128 + r, h]), 5) for r in rows[:6]], "rows", rows[:12])
print("bad runs", nbad, "of 59")
