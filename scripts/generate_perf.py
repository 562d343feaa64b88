"""generate_conditional (stage-2 decode loop with the reference's grammar rules) on a random-init Performer:
accepted events/s with the reference's reject-and-redraw semantics vs the opt-in device grammar mask."""
import sys, os, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from emo_disentanger_b200.stage2 import MusicPerformer
from emo_disentanger_b200.generate import generate_conditional
from emo_disentanger_b200.synth import synthetic_vocab, synthetic_lead_sheet

V = 329
e2i, i2e = synthetic_vocab(V, 2)
with contextlib.redirect_stdout(sys.stderr):
    m = MusicPerformer(V, 12, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128)
m = m.cuda().eval()
lead = synthetic_lead_sheet(e2i, 8, seed=1)
primer = [e2i['Emotion_Q1'], e2i['Key_C'], e2i['Tempo_110']]
for mode in (False, True):
    for rep in range(2):
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(open(os.devnull, "w")):
            toks = generate_conditional(m, e2i, i2e, lead, primer, max_events=600, temp=1.2, top_p=0.9, device_grammar=mode)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    n_new = len(toks) - len(primer) - sum(len(b) for b in lead[:1])
    print("device_grammar=%-5s: %4d events in %.3f s -> %.0f accepted events/s" % (mode, len(toks), dt, len(toks) / dt), flush=True)
