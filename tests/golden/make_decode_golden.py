"""Golden token sequences for the decode loops, produced by the REFERENCE's own loops
(`generate_conditional`, `generate_plain_xl`, extracted verbatim from /root/reference) driving the oracle
models on CPU.  Run in the build container:   python tests/golden/make_decode_golden.py

Two modes per loop:
  greedy  : the reference's `nucleus` is swapped for argmax (the reference has no greedy switch); the
            grammar checks stay ON for stage 1 and are skipped for stage 2 (`skip_check=True`, an existing
            reference flag) -> bit-exact token comparison on the GPU.
  sampled : unmodified temperature()/nucleus() with a seeded numpy global RNG -> the GPU loop, seeded the
            same way, must follow the same stream (ties on a CDF boundary excepted).
The feature map Omega is FIXED for a generation (what `omit_feature_map_draw` intends)."""
import contextlib
import io
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import performer_oracle as PO, gpt2_oracle as GO, txl_oracle as TO, ref_import  # noqa: E402
from emo_disentanger_b200.synth import synthetic_vocab, synthetic_lead_sheet  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _P(torch.nn.Module):
    """minimal nn.Module facade over an oracle forward, with the call signature the loops use"""
    def __init__(self, fn):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.fn = fn

    def forward(self, x, seg_inp=None, keep_last_only=False, attn_kwargs=None):
        return self.fn(x, seg_inp)[:, -1, :]


class _S1(torch.nn.Module):
    def __init__(self, sd, L, mem_len):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.sd, self.L, self.mem_len = sd, L, mem_len

    def generate(self, dec_input, mems):
        mems = None if (mems is None or len(mems) == 0) else mems
        return TO.txl_generate(self.sd, dec_input, mems, self.L, 8, 512, self.mem_len)


def run_quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    fns = ref_import.decode_functions()
    out = {}
    # ---------------- stage 2 ----------------
    V, L, n_bars = 96, 2, 3
    e2i, i2e = synthetic_vocab(V, 2)
    lead = synthetic_lead_sheet(e2i, n_bars, 1, events_per_bar=8)
    primer = [e2i['Emotion_Q1'], e2i['Key_C'], e2i['Tempo_110']]
    g = torch.Generator().manual_seed(5)
    omegas = [torch.randn(64, 64, generator=g) for _ in range(L)]
    sdp = PO.seeded_state(PO.performer_state_shapes(V, L), 31, std=0.05)
    sdp["pe.pe"] = PO.sinusoid_pe(12000, 512)
    sdg = PO.seeded_state(GO.gpt2_state_shapes(V, L), 32, std=0.05)
    sdg["pe.pe"] = PO.sinusoid_pe(12000, 512)
    models = {"performer": _P(lambda x, s: PO.performer_forward(sdp, x, s, omegas, L, 8, 512)),
              "gpt2": _P(lambda x, s: GO.gpt2_forward(sdg, x, s, L, 8, 512))}
    ns = fns["stage2"]
    ref_nucleus = ns["nucleus"]
    for name, mdl in models.items():
        ns["nucleus"] = lambda probs, p: int(np.argmax(probs))
        with torch.no_grad():
            toks = run_quiet(ns["generate_conditional"], mdl, e2i, i2e, [list(b) for b in lead], list(primer),
                             max_events=70, skip_check=True, temp=1.1, top_p=0.99, model_type=name)
        out["s2_%s_greedy" % name] = np.array(toks)
        ns["nucleus"] = ref_nucleus
        np.random.seed(1234)
        with torch.no_grad():
            toks = run_quiet(ns["generate_conditional"], mdl, e2i, i2e, [list(b) for b in lead], list(primer),
                             max_events=70, skip_check=False, temp=1.1, top_p=0.99, model_type=name)
        out["s2_%s_sampled" % name] = np.array(toks)
        print(name, "greedy", len(out["s2_%s_greedy" % name]), "sampled", len(out["s2_%s_sampled" % name]))
    out["s2_omegas"] = torch.stack(omegas).numpy()
    out["s2_lead"] = np.array([len(b) for b in lead] + [x for b in lead for x in b])
    # ---------------- stage 1 ----------------
    V1, L1 = 96, 2
    e2i1, i2e1 = synthetic_vocab(V1, 1)
    sd1 = PO.seeded_state(TO.txl_state_shapes(V1, L1), 33, std=0.05)
    m1 = _S1(sd1, L1, 32)
    ns1 = fns["stage1"]
    ref_nucleus1 = ns1["nucleus"]
    ns1["nucleus"] = lambda probs, p: int(np.argmax(probs))
    with torch.no_grad():
        toks, _ = run_quiet(ns1["generate_plain_xl"], m1, e2i1, i2e1, max_bars=4, max_events=48, primer=['Emotion_Positive'],
                            temp=1.2, top_p=0.97, representation='remi')
    out["s1_greedy"] = np.array(toks if toks is not None else [-1])
    ns1["nucleus"] = ref_nucleus1
    np.random.seed(4321)
    with torch.no_grad():
        res = run_quiet(ns1["generate_plain_xl"], m1, e2i1, i2e1, max_bars=4, max_events=48, primer=['Emotion_Positive'],
                        temp=1.2, top_p=0.97, representation='functional', key_determine=None)
    out["s1_sampled"] = np.array(res[0] if res[0] is not None else [-1])
    print("stage1 greedy", len(out["s1_greedy"]), "sampled", len(out["s1_sampled"]))
    np.savez_compressed(os.path.join(OUT, "decode_small.npz"), V=V, L=L, V1=V1, L1=L1, n_bars=n_bars, **out)


if __name__ == "__main__":
    main()
