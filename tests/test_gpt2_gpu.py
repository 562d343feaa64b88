"""GPU parity of the causal softmax attention kernels (K8) and of the stage-2 MusicGPT2 module against
the oracle (gpt2_oracle, pinned against the reference module + HF GPT2Block) and its golden vectors."""
import math
import numpy as np
import pytest
import torch

from helpers import golden, rel_err, rms_rel, load_seeded
from oracle import gpt2_oracle as GO, performer_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref_attn(q, k, v, scale):
    """q [B,Tq,H,64], k,v [B,Tk,H,64] (fp64) -> out [B,Tq,H,64]; query i sees keys j <= i + Tk - Tq."""
    Tq, Tk = q.shape[1], k.shape[1]
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    i = torch.arange(Tq)[:, None]
    j = torch.arange(Tk)[None, :]
    s = s.masked_fill((j > i + Tk - Tq)[None, None], float("-inf"))
    return torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), v)


@pytest.mark.parametrize("dtype,Tq,Tk", [(torch.float32, 70, 70), (torch.float32, 33, 97), (torch.bfloat16, 200, 200),
                                          (torch.bfloat16, 64, 64), (torch.bfloat16, 1, 130), (torch.bfloat16, 1, 1),
                                          # tcgen05 path (Tq >= 64): ragged tiles, Tq < Tk with an offset that is not a tile multiple
                                          (torch.bfloat16, 128, 128), (torch.bfloat16, 333, 333), (torch.bfloat16, 130, 300),
                                          (torch.bfloat16, 640, 1000), (torch.bfloat16, 1024, 1024)])
def test_attention_fwd_bwd_vs_torch(dtype, Tq, Tk):
    from emo_disentanger_b200 import ops
    B, H = 2, 8
    g = torch.Generator().manual_seed(Tq * 1000 + Tk)
    qf = torch.randn(B, Tq, H * 64, generator=g)
    kvf = torch.randn(B, Tk, 2 * H * 64, generator=g)
    qd, kvd = qf.to(DEV).to(dtype), kvf.to(DEV).to(dtype)
    q = qd.unflatten(-1, (H, 64))
    k, v = kvd[:, :, :H * 64].unflatten(-1, (H, 64)), kvd[:, :, H * 64:].unflatten(-1, (H, 64))
    out = torch.empty(B, Tq, H * 64, device=DEV, dtype=dtype)
    lse = torch.empty(B, H, Tq, device=DEV)
    ops.attn_fwd(q, k, v, out, lse, 0.125)
    qr = qd.float().cpu().double().requires_grad_(True)
    kvr = kvd.float().cpu().double().requires_grad_(True)
    ref = _ref_attn(qr.unflatten(-1, (H, 64)), kvr[:, :, :H * 64].unflatten(-1, (H, 64)),
                    kvr[:, :, H * 64:].unflatten(-1, (H, 64)), 0.125)
    tol = 1e-4 if dtype == torch.float32 else 1.5e-2
    assert rel_err(out.float().view(B, Tq, H, 64), ref.float()) < tol
    dout = torch.randn(B, Tq, H * 64, generator=g).to(dtype)
    dq = torch.empty_like(qd)
    dkv = torch.empty_like(kvd)
    ops.attn_bwd(q, k, v, out, dout.to(DEV), lse, dq.unflatten(-1, (H, 64)), dkv[:, :, :H * 64].unflatten(-1, (H, 64)),
                 dkv[:, :, H * 64:].unflatten(-1, (H, 64)), 0.125)
    ref.backward(dout.double().view(B, Tq, H, 64))
    tolb = 1e-3 if dtype == torch.float32 else 3e-2
    if Tk > 1:
        assert rms_rel(dq.float(), qr.grad.float()) < tolb
    else:                                   # single key: softmax is constant, dq == 0 exactly in exact arithmetic
        assert float(dq.float().abs().max()) < 1e-5
    assert rms_rel(dkv.float(), kvr.grad.float()) < tolb


def test_attention_dropout_mask_consistent_fwd_bwd():
    """with attention-prob dropout, backward must differentiate the SAME masked forward: check by finite
    differences of sum(out * w) along a random direction in v and q (fp32 mode)."""
    from emo_disentanger_b200 import ops
    B, H, T = 1, 8, 48
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, 3 * H * 64, generator=g).to(DEV)
    w = torch.randn(B, T, H * 64, generator=g).to(DEV)
    d = H * 64

    def fwd(xx):
        q, k, v = (xx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
        out = torch.empty(B, T, d, device=DEV)
        lse = torch.empty(B, H, T, device=DEV)
        ops.attn_fwd(q, k, v, out, lse, 0.125, 0.3, 77)
        return out, lse
    out, lse = fwd(x)
    frac0 = float((out == 0).float().mean())
    assert frac0 < 0.05                       # dropout acts on probabilities, not outputs (row 0 has one key)
    dx = torch.empty_like(x)
    q, k, v = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    dq, dk, dv = (dx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    ops.attn_bwd(q, k, v, out, w, lse, dq, dk, dv, 0.125, 0.3, 77)
    dirn = torch.randn(x.shape, generator=g).to(DEV)
    eps = 1e-2
    fp = (fwd(x + eps * dirn)[0].double() * w.double()).sum()
    fm = (fwd(x - eps * dirn)[0].double() * w.double()).sum()
    fd = float((fp - fm) / (2 * eps))
    an = float((dx.double() * dirn.double()).sum())
    assert abs(fd - an) / (abs(an) + 1e-6) < 2e-2
    # and the no-dropout output differs (the mask is really applied)
    out0 = torch.empty_like(out)
    ops.attn_fwd(q, k, v, out0, lse, 0.125)
    assert rel_err(out, out0) > 1e-2


@pytest.mark.parametrize("B,T,p", [(2, 2048, 0.0), (2, 2048, 0.1), (3, 700, 0.25), (2, 333, 0.25)])   # odd T: the per-element mask path
def test_attention_tcgen05_equals_mma_sync(B, T, p):
    """the tcgen05 + TMA attention (128 x 128 tiles, P / P^T as tensor-memory operands, fp32 dQ reductions) against the
    round-1 mma.sync kernels (64 x 64 tiles): two independent implementations of the same math and the SAME dropout
    mask (with p > 0 a different mask would show as O(1) differences)"""
    from emo_disentanger_b200 import ops, _lib
    H, d = 8, 512
    g = torch.Generator().manual_seed(100 + T)
    x = torch.randn(B, T, 3 * d, generator=g).to(DEV).to(torch.bfloat16)
    dout = torch.randn(B, T, d, generator=g).to(DEV).to(torch.bfloat16)
    q, k, v = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    res = []
    for tc in (1, 0):
        _lib.lib().emo_attn_set_tc(tc)
        try:
            out = torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16)
            lse = torch.empty(B, H, T, device=DEV)
            ops.attn_fwd(q, k, v, out, lse, 0.125, p, 4242)
            dx = torch.empty_like(x)
            dq, dk, dv = (dx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
            ops.attn_bwd(q, k, v, out, dout, lse, dq, dk, dv, 0.125, p, 4242)
            torch.cuda.synchronize()
            res.append((out.float(), lse, dx.float()))
        finally:
            _lib.lib().emo_attn_set_tc(1)
    (o1, l1, g1), (o0, l0, g0) = res
    assert torch.isfinite(o1).all() and torch.isfinite(g1).all()
    assert rms_rel(o1, o0) < 6e-3
    assert float((l1 - l0).abs().max()) < 2e-3
    for i, nm in enumerate("qkv"):
        assert rms_rel(g1[:, :, i * d:(i + 1) * d], g0[:, :, i * d:(i + 1) * d]) < 1.5e-2, "d" + nm


def test_attention_causal_prefix_invariance_full_size():
    """size-independent property at the benchmark length: outputs of the first T/2 queries do not
    depend on later keys."""
    from emo_disentanger_b200 import ops
    B, H, T = 1, 8, 2048
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, 3 * H * 64, generator=g).to(DEV).to(torch.bfloat16)
    d = H * 64
    q, k, v = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    o1 = torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=DEV)
    ops.attn_fwd(q, k, v, o1, lse, 0.125)
    x2 = x.clone()
    x2[:, T // 2:] = torch.randn(B, T // 2, 3 * d, generator=g).to(DEV).to(torch.bfloat16)
    q2, k2, v2 = (x2[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    o2 = torch.empty_like(o1)
    ops.attn_fwd(q2, k2, v2, o2, lse, 0.125)
    assert torch.equal(o1[:, :T // 2], o2[:, :T // 2])
    assert torch.isfinite(o1.float()).all()


def _model(g, dtype, dropout=0.0):
    from emo_disentanger_b200.stage2 import MusicGPT2
    V, L = int(g["V"]), int(g["L"])
    m = MusicGPT2(V, L, 8, 512, 2048, 512, dropout=dropout, use_segment_emb=True, n_segment_types=2,
                  compute_dtype=dtype)
    load_seeded(m, GO.gpt2_state_shapes(V, L), int(g["seed"]))
    return m.cuda()


def _inputs(g):
    return (torch.from_numpy(g["tok"]).cuda(), torch.from_numpy(g["seg"]).cuda(), torch.from_numpy(g["tgt"]).cuda())


def test_gpt2_fp32_logits_loss_argmax_vs_golden():
    g = golden("gpt2_small.npz")
    m = _model(g, torch.float32).eval()
    tok, seg, tgt = _inputs(g)
    with torch.no_grad():
        logits = m(tok, seg_inp=seg)
    ref = torch.from_numpy(g["logits"])
    assert rel_err(logits, ref) < 1e-3
    loss = m.compute_loss(logits, tgt)["recons_loss"]
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    am = logits.argmax(-1).cpu().numpy()
    top2 = ref.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).numpy()
    assert ((am == g["argmax"]) | (margin < 1e-4)).all()


def test_gpt2_bf16_hidden_and_logits_vs_golden():
    g = golden("gpt2_small.npz")
    m = _model(g, torch.bfloat16).eval()
    tok, seg, tgt = _inputs(g)
    with torch.no_grad():
        hid, _ = m._forward_hidden(tok, seg, save=False)
        logits = m(tok, seg_inp=seg)
    ref_h = torch.from_numpy(g["hidden_last"]).view(-1, 512)
    assert rms_rel(hid.float(), ref_h) < 1e-2
    assert rms_rel(logits, torch.from_numpy(g["logits"])) < 2e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 8e-2)])
def test_gpt2_gradients_vs_golden(dtype, tol):
    g = golden("gpt2_small.npz")
    m = _model(g, dtype).train()
    tok, seg, tgt = _inputs(g)
    m.zero_grad()
    m.compute_loss(m(tok, seg_inp=seg), tgt)["total_loss"].backward()
    named = dict(m.named_parameters())
    n = 0
    for key in g.files:
        if key.startswith("grad:"):
            e = rms_rel(named[key[5:]].grad, torch.from_numpy(g[key]))
            assert e < tol, "%s rms rel err %.3e" % (key[5:], e)
            n += 1
        elif key.startswith("gradslice:"):
            e = rms_rel(named[key[10:]].grad.reshape(-1)[:2048], torch.from_numpy(g[key]))
            assert e < tol, "%s rms rel err %.3e" % (key[10:], e)
            n += 1
    assert n >= 4


def test_gpt2_training_with_dropout_learns():
    from emo_disentanger_b200.optim import FusedAdam
    g = golden("gpt2_small.npz")
    tok, seg, tgt = _inputs(g)
    torch.manual_seed(7)
    m = _model(g, torch.bfloat16, dropout=0.1).train()
    opt = FusedAdam(m, lr=1e-3, max_grad_norm=0.5)
    losses = []
    for it in range(8):
        acc = m.train_step(tok, seg, tgt)
        opt.step()
        losses.append(float(acc[1] / acc[0]))
    assert losses[-1] < losses[0] - 0.3


def test_gpt2_loads_4_28_style_checkpoint_with_mask_buffers():
    g = golden("gpt2_small.npz")
    m = _model(g, torch.bfloat16)
    sd = dict(m.state_dict())
    for l in range(int(g["L"])):
        sd["transformer_decoder.%d.attn.bias" % l] = torch.ones(1, 1, 8, 8, dtype=torch.bool)
        sd["transformer_decoder.%d.attn.masked_bias" % l] = torch.tensor(-1e4)
    m2 = _model(g, torch.bfloat16)
    m2.load_state_dict(sd)
    assert torch.equal(m2._flat, m._flat)
