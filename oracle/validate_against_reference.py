"""Pin the oracle restatements against the UNMODIFIED reference (run in the build container):

    python -m oracle.validate_against_reference

Checks (fp32 CPU, eval mode, seeded):
  1. stage-1 PlainTransformer.forward / .generate(mems)   vs oracle.txl_oracle
  2. stage-2 MusicGPT2.forward (HF GPT2Block 5.5.0 + shim) vs oracle.gpt2_oracle
  3. stage-2 MusicPerformer.forward over the fast_transformers stand-in (glue only) vs
     oracle.performer_oracle; plus sequential vs chunked causal product, and the quadratic form
  4. temperature / nucleus (both stages) vs oracle.sampling_oracle with a seeded numpy RNG
TEST INFRASTRUCTURE; exits non-zero on mismatch."""
import sys
import numpy as np
import torch

from . import ref_import, performer_oracle as PO, gpt2_oracle as GO, txl_oracle as TO, sampling_oracle as SO


def _maxrel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def check_stage1():
    m = ref_import.stage1_model()
    torch.manual_seed(0)
    V, L = 216, 3
    model = m.PlainTransformer(512, V, L, 8, 512, 2048, 0, 64, pre_lnorm=True).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in sd:   # make biases / LN non-trivial
        if k.endswith("bias") and sd[k].dim() == 1:
            sd[k] = 0.02 * torch.randn_like(sd[k])
    model.load_state_dict(sd)
    tok = torch.randint(0, V - 1, (48, 2))
    with torch.no_grad():
        ref, _ = model(tok, tuple())
        got, _ = TO.txl_forward(sd, tok, None, L, 8, 512, 0)
    e1 = _maxrel(got, ref)
    # incremental decode with memory
    model2 = m.PlainTransformer(512, V, L, 8, 512, 2048, 16, 16, pre_lnorm=True).eval()
    model2.load_state_dict(sd)
    mems_r, mems_o, e2 = tuple(), None, 0.0
    with torch.no_grad():
        for step in range(24):
            inp = tok[:5, :1] if step == 0 else tok[5 + step:6 + step, :1]
            lr, mems_r = model2.generate(inp, mems_r)
            lo, mems_o = TO.txl_generate(sd, inp, mems_o, L, 8, 512, 16)
            e2 = max(e2, _maxrel(lo, lr))
            assert all(a.shape == b.shape for a, b in zip(mems_r, mems_o))
    print("stage1 forward rel err %.2e, generate(mems) rel err %.2e" % (e1, e2))
    return e1 < 1e-5 and e2 < 1e-5


def check_gpt2():
    m = ref_import.stage2_gpt2()
    torch.manual_seed(1)
    V, L = 372, 2
    model = m.MusicGPT2(V, L, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in sd:
        if k.endswith("bias") and sd[k].dim() == 1:
            sd[k] = 0.02 * torch.randn_like(sd[k])
    model.load_state_dict(sd)
    tok = torch.randint(0, V - 1, (2, 96)); seg = torch.randint(0, 2, (2, 96))
    with torch.no_grad():
        ref = model(tok, seg_inp=seg)
        got = GO.gpt2_forward(sd, tok, seg, L, 8, 512)
    e = _maxrel(got, ref)
    print("gpt2 forward rel err %.2e" % e)
    return e < 1e-5


def check_performer():
    m = ref_import.stage2_performer()
    torch.manual_seed(2)
    V, L = 329, 2
    model = m.MusicPerformer(V, L, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2,
                             favor_feature_dims=128).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k in sd:
        if k.endswith("bias") and sd[k].dim() == 1:
            sd[k] = 0.02 * torch.randn_like(sd[k])
    model.load_state_dict(sd)
    omegas = [PO.draw_omega(64, 64) for _ in range(L)]
    for l, layer in enumerate(model.transformer_decoder.decoder_layers):
        layer.attention.inner_attention.feature_map.injected = omegas[l]
    tok = torch.randint(0, V - 1, (2, 80)); seg = torch.randint(0, 2, (2, 80))
    with torch.no_grad():
        ref = model(tok, seg_inp=seg)
        got = PO.performer_forward(sd, tok, seg, omegas, L, 8, 512)
    e = _maxrel(got, ref)
    # causal product: sequential (cpp statement) vs chunked vs explicit quadratic form (fp64)
    q, k, v = (torch.randn(1, 40, 2, 64, dtype=torch.float64) for _ in range(3))
    om = PO.draw_omega(64, 64, dtype=torch.float64)
    o_seq, _ = PO.causal_linear_attention(q, k, v, om, sequential=True)
    o_chk, _ = PO.causal_linear_attention(q, k, v, om, sequential=False)
    Q, K = PO.favor_features(q, om), PO.favor_features(k, om)
    A = torch.einsum("nlhi,njhi->nhlj", Q, K) * torch.tril(torch.ones(40, 40, dtype=torch.float64))
    o_quad = torch.einsum("nhlj,njhd->nlhd", A, v) / (A.sum(-1).permute(0, 2, 1)[..., None] + 1e-6)
    e2 = max(_maxrel(o_seq, o_quad), _maxrel(o_chk, o_quad))
    print("performer glue rel err %.2e, causal product forms rel err %.2e" % (e, e2))
    return e < 1e-5 and e2 < 1e-10


def check_sampling():
    fns = ref_import.sampling_functions()
    ok = True
    rng = np.random.RandomState(7)
    n = 0
    for stage, (temperature, nucleus) in fns.items():
        for trial in range(200):
            V = int(rng.choice([216, 329, 372]))
            logits = (rng.randn(V) * rng.choice([0.5, 2.0, 6.0])).astype(np.float32)
            t = float(rng.choice([1.0, 1.1, 1.2])); p = float(rng.choice([0.9, 0.97, 0.99]))
            seed = int(rng.randint(1 << 30))
            try:
                np.random.seed(seed)
                kw = {"inadmissibles": None} if stage == "stage2" else {}
                ref = int(nucleus(temperature(logits.copy(), t, **kw), p))
            except IndexError:
                ref = "IndexError"
            try:
                u = np.random.RandomState(seed).random_sample()
                got = SO.sample(logits.copy(), t, p, u)
            except IndexError:
                got = "IndexError"
            ok &= (ref == got)
            n += 1
    print("sampling: %d cases, all equal = %s" % (n, ok))
    return ok


def main():
    if not ref_import.available():
        print("reference not present; nothing to validate")
        return 0
    torch.set_grad_enabled(False)
    res = [check_stage1(), check_gpt2(), check_performer(), check_sampling()]
    print("ALL OK" if all(res) else "MISMATCH")
    return 0 if all(res) else 1


if __name__ == "__main__":
    sys.exit(main())
