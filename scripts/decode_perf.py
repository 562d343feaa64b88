"""Time the stage-2 decode step (model step + device sampler + the one-int D2H) at batch 1 / 4."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib
import numpy as np
import torch
from emo_disentanger_b200.stage2 import MusicPerformer, MusicGPT2
from emo_disentanger_b200.decode import Stage2Decoder
from emo_disentanger_b200.generate import DeviceSampler

V = 329
N = int(os.environ.get("N", 256))
torch.manual_seed(0)
np.random.seed(0)
for kind in sys.argv[1:] or ["performer", "gpt2"]:
    with contextlib.redirect_stdout(sys.stderr):
        if kind == "performer":
            m = MusicPerformer(V, 12, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128)
        else:
            m = MusicGPT2(V, 12, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2)
    m = m.cuda().eval()
    for B in (1, 4):
        dec = Stage2Decoder(m, batch=B, max_len=2048, use_pdl=bool(int(os.environ.get("PDL", "0"))))
        smp = DeviceSampler(dec.dev, rows=B)
        for b in range(B):
            dec.append(b, list(range(3, 40)), [0] * 37)
        toks = [5] * B
        for phase in ("warm", "timed"):
            n = 16 if phase == "warm" else N
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n):
                if kind == "performer":
                    toks, _ = dec.step_sample(toks, [1] * B, np.random.random_sample(B), 1.2, 0.9)
                else:
                    lg = dec.step(toks, [1] * B)
                    toks = smp.draw(lg, V, 1.2, 0.9)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        # model step alone (no sampler / no D2H)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(N):
            dec.step(toks, [1] * B)
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t0
        print("%-9s B=%d: %7.1f us/step with sampler+D2H (%8.0f tok/s) | model step only %7.1f us" %
              (kind, B, dt / N * 1e6, B * N / dt, dt2 / N * 1e6), flush=True)
