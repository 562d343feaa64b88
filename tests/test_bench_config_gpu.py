"""Parity AT the benchmark configuration (BASELINE.json configs[1]: stage-2 Performer d=512 / 8 heads / 12 layers,
functional vocabulary V = 329, T = 2048): the CUDA path against the CPU oracle on the same seeded inputs.

The GPU side runs the bench's per-GPU batch (74 sequences: the multi-wave FAVOR+ plan, 592 (batch, head) items on
148 SMs; CTA-pair GEMM tiles) with the oracle's sequences planted in its first and last rows; the oracle side runs
those two sequences (B = 1 each, a few seconds of CPU).  Tolerances are the north star's: 1e-3 rel on fp32 logits,
1e-2 on bf16 hidden states."""
import pytest
import torch

from helpers import rel_err, rms_rel, load_seeded
from oracle import performer_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"
V, L, T, H, BB = 329, 12, 2048, 8, 74


def _model(dtype, seed=11):
    from emo_disentanger_b200.stage2 import MusicPerformer
    m = MusicPerformer(V, L, H, 512, 2048, 512, dropout=0.0, use_segment_emb=True, n_segment_types=2,
                       favor_feature_dims=128, compute_dtype=dtype)
    sd = load_seeded(m, PO.performer_state_shapes(V, L), seed)
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(seed + 1)
    om = torch.randn(L, 64, 64, generator=gen)
    m.fixed_omegas = om.cuda()
    tok = torch.randint(0, V - 1, (BB, T), generator=gen)
    seg = torch.randint(0, 2, (BB, T), generator=gen)
    return m, sd, om, tok, seg


def _oracle_rows(sd, om, tok, seg, rows, taps_out=None):
    outs = []
    for r in rows:
        taps = []
        with torch.no_grad():
            outs.append(PO.performer_forward(sd, tok[r:r + 1], seg[r:r + 1], [om[l] for l in range(L)], L, H, 512, taps=taps))
        if taps_out is not None:
            taps_out.append(taps[-1])            # output of the last layer = the final hidden state
    return torch.cat(outs, 0)


def test_fp32_logits_at_bench_config_vs_oracle():
    m, sd, om, tok, seg = _model(torch.float32)
    rows = [0, BB - 1]
    ref = _oracle_rows(sd, om, tok, seg, rows)
    B = 8                                                     # fp32 parity mode: SIMT GEMMs, keep the batch small
    idx = [0] * (B - 1) + [BB - 1]
    idx[1:B - 1] = range(1, B - 1)
    with torch.no_grad():
        out = m(tok[idx].cuda(), seg_inp=seg[idx].cuda())
    got = torch.stack([out[0], out[B - 1]]).float().cpu()
    assert rel_err(got, ref) < 1e-3                           # north-star tolerance on fp32 logits
    top2 = ref.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    assert ((got.argmax(-1) == ref.argmax(-1)) | (margin < 1e-4)).all()     # greedy tokens identical (ties excepted)


def test_bf16_hidden_and_logits_at_bench_config_vs_oracle():
    m, sd, om, tok, seg = _model(torch.bfloat16)
    rows = [0, BB - 1]
    ref_h = []
    ref = _oracle_rows(sd, om, tok, seg, rows, ref_h)
    with torch.no_grad():
        hid, _ = m._forward_hidden(tok.cuda(), seg.cuda(), save=False)
        out = m(tok.cuda(), seg_inp=seg.cuda())
    got = torch.stack([out[0], out[BB - 1]]).float().cpu()
    assert torch.isfinite(out.float()).all()
    assert rms_rel(got, ref) < 2e-2
    hid = hid.view(BB, T, 512)
    got_h = torch.stack([hid[0], hid[BB - 1]]).float().cpu()
    e = rms_rel(got_h, torch.cat(ref_h, 0))
    assert e < 1e-2, "bf16 hidden states after 12 layers: rms rel err %.3e" % e     # north-star tolerance
    # rows that share a sequence share the result, wherever they sit in the batch (multi-wave / multi-segment plan)
    tok2, seg2 = tok.clone(), seg.clone()
    tok2[37], seg2[37] = tok[0], seg[0]
    with torch.no_grad():
        out2 = m(tok2.cuda(), seg_inp=seg2.cuda())
    assert rms_rel(out2[37].float(), out[0].float()) < 1e-3


@pytest.fixture(scope="module")
def ops():
    from emo_disentanger_b200 import ops as o
    return o


@pytest.mark.parametrize("B", [BB, 4])       # 74 = the bench batch (one segment per (b, h)); 4 = the reference batch_size (segmented plan)
def test_favor_fwd_bwd_at_bench_length_vs_oracle_autograd(ops, B):
    """one layer's attention core at T = 2048: out, den, dq, dk, dv against the oracle and its autograd"""
    g = torch.Generator().manual_seed(5 + B)
    d = H * 64
    qkv = (torch.randn(B, T, 3 * d, generator=g) * 0.7).to(torch.bfloat16)
    omega = torch.randn(64, 64, generator=g)
    dout = torch.randn(B, T, d, generator=g).to(torch.bfloat16)
    qkv_d = qkv.to(DEV)
    q, k, v = (qkv_d[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    out = torch.empty(B, T, d, device=DEV, dtype=torch.bfloat16)
    den = torch.empty(B, T, H, device=DEV)
    ws = ops.favor_workspace(B, T, H, torch.bfloat16, DEV)
    ops.favor_fwd(q, k, v, omega.to(DEV), out, den, seg_states=ws)
    dqkv = torch.empty_like(qkv_d)
    dq, dk, dv = (dqkv[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    ops.favor_bwd(q, k, v, omega.to(DEV), out, dout.to(DEV), den, ws, dq, dk, dv)
    for r in (0, B - 1):
        x = qkv[r:r + 1].double().requires_grad_(True)
        qo, ko, vo = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
        ref, rden = PO.causal_linear_attention(qo, ko, vo, omega.double())
        ref.backward(dout[r:r + 1].double().view(1, T, H, 64))
        assert rms_rel(out[r].float().cpu().view(T, H, 64), ref[0].float()) < 1.5e-2
        assert rms_rel(den[r].cpu(), rden[0].float()) < 2e-2
        for i, name in enumerate("qkv"):
            e = rms_rel(dqkv[r, :, i * d:(i + 1) * d].float().cpu(), x.grad[0, :, i * d:(i + 1) * d].float())
            assert e < 3e-2, "d%s rms rel err %.3e (row %d)" % (name, e, r)


def test_external_optimizer_and_reload_reach_the_bf16_gemm_weights():
    """torch.optim.Adam.step() and load_state_dict() after a forward must change the next forward
    (the bf16 shadow of the flat fp32 masters used to be refreshed from a version counter that never moved)."""
    from emo_disentanger_b200.stage2 import MusicPerformer
    m = MusicPerformer(50, 1, 8, 512, 2048, 512, dropout=0.0, use_segment_emb=True, n_segment_types=2,
                       favor_feature_dims=128).cuda().train()
    m.fixed_omegas = torch.randn(1, 64, 64, generator=torch.Generator().manual_seed(0)).cuda()
    tok = torch.randint(0, 49, (2, 64), device=DEV)
    seg = torch.randint(0, 2, (2, 64), device=DEV)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    l0 = m(tok, seg_inp=seg)
    loss = m.compute_loss(l0, tok)["total_loss"]
    loss.backward()
    w_name = "transformer_decoder.decoder_layers.0.linear1.weight"
    w_before = dict(m.named_parameters())[w_name].detach().clone()
    opt.step()
    assert not torch.equal(dict(m.named_parameters())[w_name], w_before)
    with torch.no_grad():
        l1 = m(tok, seg_inp=seg)
    off, n, shp = m._slices()[w_name]
    assert torch.equal(m._flat_lp[off:off + n].view(shp), dict(m.named_parameters())[w_name].to(torch.bfloat16))
    assert float((l1 - l0).abs().max()) > 1e-3
    sd = {k: (v * 0.5 if v.dtype.is_floating_point and "pe.pe" not in k else v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    with torch.no_grad():
        l2 = m(tok, seg_inp=seg)
    assert float((l2 - l1).abs().max()) > 1e-3
    assert torch.equal(m._flat_lp[off:off + n].view(shp), dict(m.named_parameters())[w_name].to(torch.bfloat16))


def test_orthogonal_feature_draw():
    """SURVEY A5 / App. A.1 reads fast-transformers' Favor as always drawing Omega through a QR (orthogonal random
    features); the oracle (performer_oracle.draw_omega) reads the default as a plain Gaussian draw and QR only with
    orthogonal=True.  Both are implemented (MusicPerformer(orthogonal_features=...)); DESIGN.md section 4 records the
    disagreement.  The orthogonal draw: columns mutually orthogonal, column j scaled by the norm of row j of the
    Gaussian block."""
    from emo_disentanger_b200.stage2 import MusicPerformer
    m = MusicPerformer(50, 2, 8, 512, 2048, 512, favor_feature_dims=128, orthogonal_features=True)
    torch.manual_seed(3)
    om = m.draw_omegas("cuda").cpu().double()
    assert om.shape == (2, 64, 64)
    for l in range(2):
        gram = om[l].t() @ om[l]
        off = gram - torch.diag(torch.diag(gram))
        assert float(off.abs().max()) < 1e-3 * float(torch.diag(gram).mean())
        n = torch.diag(gram).sqrt()                            # chi(64)-distributed column norms: mean ~ 7.97
        assert 7.0 < float(n.mean()) < 9.0 and float(n.std()) > 0.3
    # the oracle's orthogonal draw has the same structure
    o = PO.draw_omega(64, 64, generator=torch.Generator().manual_seed(1), orthogonal=True).double()
    gram = o.t() @ o
    assert float((gram - torch.diag(torch.diag(gram))).abs().max()) < 1e-3 * float(torch.diag(gram).mean())
    m2 = MusicPerformer(50, 1, 8, 512, 2048, 512, favor_feature_dims=128)       # reference call site: no kwarg -> Gaussian
    g2 = m2.draw_omegas("cuda").cpu().double()[0]
    gram2 = g2.t() @ g2
    assert float((gram2 - torch.diag(torch.diag(gram2))).abs().max()) > 0.1 * float(torch.diag(gram2).mean())
