"""A few un-graphed decode steps (for an ncu launch list of the per-token kernel sequence)."""
import sys, os, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200.stage2 import MusicPerformer
from emo_disentanger_b200.decode import Stage2Decoder
with contextlib.redirect_stdout(sys.stderr):
    m = MusicPerformer(329, 12, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128)
m = m.cuda().eval()
B = int(os.environ.get("B", 1))
dec = Stage2Decoder(m, batch=B, max_len=2048, use_graph=False)
for b in range(B):
    dec.append(b, list(range(3, 20)), [0] * 17)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(3):
    dec.step([5] * B, [1] * B)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
