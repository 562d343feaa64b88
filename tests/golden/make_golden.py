"""Generate the committed golden fixtures (run in the build container, where /root/reference
exists):   python tests/golden/make_golden.py

Each fixture = seeded synthetic inputs + outputs of the ORACLE restatement, after the script has
re-checked the oracle against the unmodified reference modules (oracle.validate_against_reference).
Weights are not stored (too large): they are regenerated from `seeded_state(shapes, seed)` with the
CPU generator, and a checksum of them is stored to detect RNG drift."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import performer_oracle as PO, gpt2_oracle as GO, txl_oracle as TO, sampling_oracle as SO  # noqa: E402
from oracle import validate_against_reference as VR, ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def synth_batch(V, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, V - 1, (B, T), generator=g)
    seg = torch.randint(0, 2, (B, T), generator=g)
    tgt = torch.roll(tok, -1, dims=1)
    tgt = torch.where(seg == 1, tgt, torch.full_like(tgt, V - 1))      # loss only on the "Full" track
    return tok, seg, tgt


def grads_of(loss, sd, keys):
    gs = torch.autograd.grad(loss, [sd[k] for k in keys])
    return {k: g for k, g in zip(keys, gs)}


def performer():
    V, L, B, T = 329, 2, 2, 96
    shapes = PO.performer_state_shapes(V, L)
    sd = PO.seeded_state(shapes, seed=11)
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    for k in shapes:
        sd[k].requires_grad_(True)
    tok, seg, tgt = synth_batch(V, B, T, 5)
    g = torch.Generator().manual_seed(3)
    omegas = [PO.draw_omega(64, 64, generator=g) for _ in range(L)]
    taps = []
    logits = PO.performer_forward(sd, tok, seg, omegas, L, 8, 512, taps=taps)
    loss = PO.ce_loss(logits, tgt, V - 1)
    keys = ["dec_out_proj.bias", "token_emb.emb_lookup.weight", "segemb.emb_lookup.weight",
            "transformer_decoder.decoder_layers.0.attention.query_projection.bias",
            "transformer_decoder.decoder_layers.0.attention.value_projection.bias",
            "transformer_decoder.decoder_layers.0.norm1.weight",
            "transformer_decoder.decoder_layers.1.linear1.bias",
            "transformer_decoder.decoder_layers.1.norm2.bias",
            "transformer_decoder.decoder_layers.0.attention.key_projection.weight"]
    gr = grads_of(loss, sd, keys)
    np.savez_compressed(
        os.path.join(OUT, "performer_small.npz"), V=V, L=L, seed=11, wsum=checksum({k: sd[k].detach() for k in shapes}),
        tok=tok.numpy(), seg=seg.numpy(), tgt=tgt.numpy(), omegas=torch.stack(omegas).numpy(),
        logits=logits.detach().numpy(), loss=float(loss), hidden_last=taps[-1].detach().numpy(),
        argmax=logits.detach().argmax(-1).numpy(),
        **{"grad:" + k: v.numpy() for k, v in gr.items() if v.numel() <= 4096},
        **{"gradnorm:" + k: float(v.norm()) for k, v in gr.items()},
        **{"gradslice:" + k: v.reshape(-1)[:2048].numpy() for k, v in gr.items() if v.numel() > 4096})
    print("performer golden: loss %.6f" % float(loss))


def gpt2():
    V, L, B, T = 372, 2, 2, 96
    shapes = GO.gpt2_state_shapes(V, L)
    sd = PO.seeded_state(shapes, seed=12)
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    for k in shapes:
        sd[k].requires_grad_(True)
    tok, seg, tgt = synth_batch(V, B, T, 6)
    taps = []
    logits = GO.gpt2_forward(sd, tok, seg, L, 8, 512, taps=taps)
    loss = PO.ce_loss(logits, tgt, V - 1)
    keys = ["dec_out_proj.bias", "transformer_decoder.0.attn.c_attn.bias", "transformer_decoder.0.ln_1.weight",
            "transformer_decoder.1.mlp.c_fc.bias", "transformer_decoder.1.attn.c_proj.weight"]
    gr = grads_of(loss, sd, keys)
    np.savez_compressed(
        os.path.join(OUT, "gpt2_small.npz"), V=V, L=L, seed=12, wsum=checksum({k: sd[k].detach() for k in shapes}),
        tok=tok.numpy(), seg=seg.numpy(), tgt=tgt.numpy(), logits=logits.detach().numpy(), loss=float(loss),
        hidden_last=taps[-1].detach().numpy(), argmax=logits.detach().argmax(-1).numpy(),
        **{"grad:" + k: v.numpy() for k, v in gr.items() if v.numel() <= 4096},
        **{"gradnorm:" + k: float(v.norm()) for k, v in gr.items()},
        **{"gradslice:" + k: v.reshape(-1)[:2048].numpy() for k, v in gr.items() if v.numel() > 4096})
    print("gpt2 golden: loss %.6f" % float(loss))


def txl():
    V, L, B, T = 216, 2, 2, 48
    shapes = TO.txl_state_shapes(V, L)
    sd = PO.seeded_state(shapes, seed=13)
    for k in shapes:
        sd[k].requires_grad_(True)
    g = torch.Generator().manual_seed(7)
    tok = torch.randint(0, V - 1, (T, B), generator=g)
    tgt = torch.roll(tok, -1, dims=0)
    tgt[-5:, 1] = V - 1
    logits, _ = TO.txl_forward(sd, tok, None, L, 8, 512, 0)
    loss = PO.ce_loss(logits, tgt, V - 1)
    keys = ["dec_out_proj.bias", "decoder.r_w_bias", "decoder.r_r_bias", "decoder.layers.0.dec_attn.layer_norm.weight",
            "decoder.layers.1.pos_ff.CoreNet.0.bias", "decoder.layers.0.dec_attn.r_net.weight",
            "decoder.layers.0.dec_attn.qkv_net.weight"]
    gr = grads_of(loss, sd, keys)
    # incremental decode with memory (mem_len 16): primer of 5 tokens then 20 single steps
    sdd = {k: v.detach() for k, v in sd.items()}
    mems, dec_logits = None, []
    with torch.no_grad():
        for step in range(21):
            inp = tok[:5, :1] if step == 0 else tok[4 + step:5 + step, :1]
            lg, mems = TO.txl_generate(sdd, inp, mems, L, 8, 512, 16)
            dec_logits.append(lg)
    np.savez_compressed(
        os.path.join(OUT, "txl_small.npz"), V=V, L=L, seed=13, wsum=checksum({k: sd[k].detach() for k in shapes}),
        tok=tok.numpy(), tgt=tgt.numpy(), logits=logits.detach().numpy(), loss=float(loss),
        argmax=logits.detach().argmax(-1).numpy(), dec_logits=torch.stack(dec_logits).numpy(),
        **{"grad:" + k: v.numpy() for k, v in gr.items() if v.numel() <= 4096},
        **{"gradnorm:" + k: float(v.norm()) for k, v in gr.items()},
        **{"gradslice:" + k: v.reshape(-1)[:2048].numpy() for k, v in gr.items() if v.numel() > 4096})
    print("txl golden: loss %.6f" % float(loss))


def sampling():
    """Golden vectors produced by the REFERENCE temperature()/nucleus() themselves."""
    fns = ref_import.sampling_functions()
    temperature, nucleus = fns["stage2"]
    rng = np.random.RandomState(123)
    cases = []
    for i in range(64):
        V = int(rng.choice([216, 329, 372]))
        logits = (rng.randn(V) * rng.choice([0.5, 2.0, 6.0])).astype(np.float32)
        t = float(rng.choice([1.0, 1.1, 1.2]))
        p = float(rng.choice([0.9, 0.97, 0.99]))
        seed = int(rng.randint(1 << 30))
        np.random.seed(seed)
        try:
            word = int(nucleus(temperature(logits.copy(), t, inadmissibles=None), p))
        except IndexError:
            word = -1
        u = np.random.RandomState(seed).random_sample()
        cases.append((logits, t, p, u, word))
    np.savez_compressed(os.path.join(OUT, "sampling_ref.npz"),
                        logits=np.array([np.pad(c[0], (0, 372 - len(c[0])), constant_values=np.nan) for c in cases]),
                        V=np.array([len(c[0]) for c in cases]), t=np.array([c[1] for c in cases]),
                        p=np.array([c[2] for c in cases]), u=np.array([c[3] for c in cases]),
                        word=np.array([c[4] for c in cases]))
    print("sampling golden: %d cases" % len(cases))


if __name__ == "__main__":
    if ref_import.available():
        assert VR.main() == 0, "oracle does not match the reference; refusing to write goldens"
    torch.set_grad_enabled(True)
    performer(); gpt2(); txl(); sampling()
