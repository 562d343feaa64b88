"""Train-step throughput of the other two models of the path (BASELINE configs[2] and the GPU twin of configs[0]):
stage-2 GPT-2 (REMI V=372, T=2048) and stage-1 PlainTransformer (functional V=216, T=512), bf16, synthetic data."""
import sys, os, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from emo_disentanger_b200.stage2 import MusicGPT2, MusicPerformer
from emo_disentanger_b200.stage1 import PlainTransformer
from emo_disentanger_b200.optim import FusedAdam
from emo_disentanger_b200.synth import synthetic_batch

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for kind in sys.argv[1:] or ["gpt2", "stage1"]:
    with contextlib.redirect_stdout(sys.stderr):
        if kind == "gpt2":
            V, T, B = 372, 2048, int(os.environ.get("B", 16))
            m = MusicGPT2(V, 12, 8, 512, 2048, 512, dropout=0.1, use_segment_emb=True, n_segment_types=2)
        else:
            V, T, B = 216, 512, int(os.environ.get("B", 64))
            m = PlainTransformer(512, V, 12, 8, 512, 2048, 0, T, dec_dropout=0.1, pre_lnorm=True)
    m = m.cuda().train()
    opt = FusedAdam(m, lr=1e-4, max_grad_norm=0.5)
    tok, seg, tgt = (t.cuda() for t in synthetic_batch(V, B, T, 0))
    if kind == "gpt2":
        step = lambda: (m.train_step(tok, seg, tgt), opt.step())
    else:
        ti, tt = tok.t().contiguous(), tgt.t().contiguous()
        step = lambda: (m.train_step(ti, tt), opt.step())
    ms = timeit(step, 8)
    print("%-7s B=%d T=%d V=%d: %.2f ms/step  %.0f tokens/s" % (kind, B, T, V, ms, B * T / ms * 1e3), flush=True)
