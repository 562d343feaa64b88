"""one warm + one profiled train step of the GPT-2 / stage-1 bench configurations (target of an ncu launch list)"""
import sys, os, contextlib, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
which = sys.argv[1]
from emo_disentanger_b200.optim import FusedAdam
if which == "gpt2":
    from emo_disentanger_b200.stage2 import MusicGPT2
    from emo_disentanger_b200.synth import synthetic_batch
    with contextlib.redirect_stdout(sys.stderr):
        m = MusicGPT2(372, 12, 8, 512, 2048, 512, dropout=0.1, use_segment_emb=True, n_segment_types=2).cuda().train()
    tok, seg, tgt = (t.cuda() for t in synthetic_batch(372, 48, 2048, 0))
    step = lambda: m.train_step(tok, seg, tgt)
else:
    from emo_disentanger_b200.stage1 import PlainTransformer
    with contextlib.redirect_stdout(sys.stderr):
        m = PlainTransformer(512, 216, 12, 8, 512, 2048, 0, 512, dec_dropout=0.1, pre_lnorm=True).cuda().train()
    tok = torch.randint(0, 215, (512, 64)).cuda()
    tgt = torch.roll(tok, -1, 0)
    step = lambda: m.train_step(tok, tgt)
opt = FusedAdam(m, lr=1e-4, max_grad_norm=0.5)
for _ in range(3):
    step(); opt.step()
torch.cuda.synchronize()
