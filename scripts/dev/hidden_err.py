"""bf16 hidden-state error per layer at the bench configuration (dev aid for the 1e-2 north-star tolerance)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from helpers import rms_rel, load_seeded
from oracle import performer_oracle as PO
import test_bench_config_gpu as TB
m, sd, om, tok, seg = TB._model(torch.bfloat16)
taps = []
with torch.no_grad():
    PO.performer_forward(sd, tok[0:1], seg[0:1], [om[l] for l in range(TB.L)], TB.L, TB.H, 512, taps=taps)
    hid, _ = m._forward_hidden(tok.cuda(), seg.cuda(), save=False)
print("taps", len(taps), "final rms rel %.4e" % rms_rel(hid.view(TB.BB, TB.T, 512)[0:1].float().cpu(), taps[-1]))
