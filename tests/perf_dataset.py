"""Stage-2 batch assembly: GPU token store (one launch per batch) vs the reference algorithm on the host (oracle
port of REMISkylineToMidiTransformerDataset.__getitem__, one item at a time as the DataLoader workers do)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root (oracle/ is test infrastructure: this timing script lives under tests/)
import numpy as np
import torch
from emo_disentanger_b200.data import Stage2TokenStore
from emo_disentanger_b200.synth import synthetic_vocab
from oracle import dataset_oracle as DO

V, T, B = 329, int(os.environ.get("T", 3072)), int(os.environ.get("B", 74))
e2i, _ = synthetic_vocab(V, 2)
e2i = {k: v for k, v in e2i.items() if k != 'PAD_None'}
i2e = {v: k for k, v in e2i.items()}
rng = np.random.RandomState(0)
pieces = [DO.synthetic_piece(e2i, int(rng.randint(30, 120)), rng, lead_len=(5, 30), full_len=(20, 200)) for _ in range(256)]
st = Stage2TokenStore(pieces, e2i, i2e, model_dec_seqlen=T, device="cuda")
print("store: %d pieces, %.1f M tokens, %.1f MB in HBM" % (len(st), st.tokens.numel() / 1e6, st.tokens.numel() * 4 / 1e6))
idx = [int(i) for i in rng.randint(0, len(st), B)]
bars = [st.piece_admissible_stbars[i][0] for i in idx]
for _ in range(3):
    st.batch(idx, bars)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n):
    out = st.batch(idx, bars)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
byts = B * T * 5 * 8
print("GPU  : %8.1f us per batch of %d x %d  (%.2f M samples/s, %.1f G tokens/s; %.0f GB/s of the 40 B/token written)"
      % (us, B, T, B / us, B * T / us / 1e3, byts / us / 1e3))
# the launch alone (selection already on the device)
from emo_disentanger_b200 import _lib as L
sel, o, ln = st.last_launch
P = lambda t: t.data_ptr()
def launch():
    L.check(L.lib().emo_stage2_batch(P(st.tokens), P(st.piece_off), P(st.bar_off), P(st.mel_start), P(st.ch_start), P(st.ch_end),
                                     P(st.flags), P(sel[0]), P(sel[1]), P(o[0]), P(o[1]), P(o[2]), P(o[3]), P(o[4]), P(ln), B, T,
                                     st.pad_token, st.eos_token, 0, torch.cuda.current_stream().cuda_stream))
for _ in range(3): launch()
torch.cuda.synchronize()
e0.record()
for _ in range(n): launch()
e1.record(); torch.cuda.synchronize()
ku = e0.elapsed_time(e1) / n * 1e3
print("kernel: %7.1f us per launch  -> %.0f GB/s of the 40 B/token written (+ 4 B/token read)" % (ku, byts / ku / 1e3))
is_chord, is_note = DO.vocab_flags(i2e, st.pad_token)
toks = [[e2i[e] for e in p[2]] for p in pieces]
t0 = time.perf_counter()
for i, b in zip(idx, bars):
    DO.assemble(toks[i], pieces[i][0], pieces[i][1], b, T, st.pad_token, st.eos_token, is_chord, is_note)
dt = time.perf_counter() - t0
print("host : %8.1f us per batch (oracle port, 1 core, events already converted to ids; the reference additionally unpickles "
      "the piece and maps strings per item)  -> %.0fx" % (dt * 1e6, dt * 1e6 / us))

# ---- stage 1 (lead sheets): Stage1TokenStore vs the oracle port of SkylineFullSongTransformerDataset + collate_fn ----
from emo_disentanger_b200.data import Stage1TokenStore, formats as F
T1, B1 = int(os.environ.get("T1", 2400)), int(os.environ.get("B1", 64))
rng = np.random.RandomState(1)
p1 = [DO.synthetic_stage1_piece(rng, int(rng.randint(20, 160)), bar_len=(10, 40)) for _ in range(256)]
e2, i2 = F.build_dictionary([ev for _, ev in p1], relative=True, **F.VOCAB_FLAGS["stage1_lead_sheet"])
s1 = Stage1TokenStore(p1, e2, i2, model_dec_seqlen=T1, device="cuda")
idx = [int(i) for i in rng.randint(0, len(s1), B1)]
for _ in range(3):
    s1.batch(idx)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    s1.batch(idx)
e1.record(); torch.cuda.synchronize()
us1 = e0.elapsed_time(e1) / n * 1e3
print("stage-1 GPU : %8.1f us per batch of %d x %d (%.0f GB/s of the 32 B/token written)" % (us1, B1, T1, B1 * T1 * 32 / us1 / 1e3))
ic, inn = DO.vocab_flags(i2, s1.pad_token)
tk = [[e2[F.event_name(e)] for e in ev] for _, ev in p1]
t0 = time.perf_counter()
for i in idx:
    DO.stage1_assemble(tk[i], p1[i][0], T1, 192, s1.pad_token, s1.eos_token, s1.bar_token, ic, inn)
dt = time.perf_counter() - t0
print("stage-1 host: %8.1f us per batch (oracle port, 1 core, ids pre-converted, decoder side only; the reference also unpickles, "
      "deep-copies and builds the unused encoder features per item)  -> %.0fx" % (dt * 1e6, dt * 1e6 / us1))
