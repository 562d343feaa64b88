"""host-side profile (cProfile) of the Performer train step at the reference's batch size 4: where the Python time goes"""
import sys, os, cProfile, pstats, contextlib, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch
from emo_disentanger_b200.stage2 import MusicPerformer
from emo_disentanger_b200.optim import FusedAdam
from emo_disentanger_b200.synth import synthetic_batch
with contextlib.redirect_stdout(sys.stderr):
    m = MusicPerformer(329, 12, 8, 512, 2048, 512, dropout=0.1, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128).cuda().train()
opt = FusedAdam(m, lr=1e-4, max_grad_norm=0.5)
tok, seg, tgt = (t.cuda() for t in synthetic_batch(329, int(os.environ.get("B", 4)), 2048, 0))
def step():
    m.train_step(tok, seg, tgt); opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("ms/step %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
t0 = time.perf_counter()
for _ in range(20): step()
host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
print("host ms/step (no sync) %.3f" % host)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
