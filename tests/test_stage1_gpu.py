"""GPU parity of the stage-1 PlainTransformer (rel-pos attention kernels K9 + the module) against the
oracle (txl_oracle, pinned against the UNMODIFIED reference module) and its golden vectors."""
import numpy as np
import pytest
import torch

from helpers import golden, rel_err, rms_rel, load_seeded
from oracle import txl_oracle as TO, performer_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(g, dtype, mem_len=0, dropout=0.0):
    from emo_disentanger_b200.stage1 import PlainTransformer
    V, L = int(g["V"]), int(g["L"])
    m = PlainTransformer(512, V, L, 8, 512, 2048, mem_len, 512, dec_dropout=dropout, pre_lnorm=True, compute_dtype=dtype)
    load_seeded(m, TO.txl_state_shapes(V, L), int(g["seed"]))
    return m.cuda()


def test_stage1_fp32_logits_loss_argmax_vs_golden():
    g = golden("txl_small.npz")
    m = _model(g, torch.float32).eval()
    tok, tgt = torch.from_numpy(g["tok"]).cuda(), torch.from_numpy(g["tgt"]).cuda()
    with torch.no_grad():
        logits, mems = m(tok, tuple())
    assert mems is None                                   # mem_len = 0: no recurrence in training
    assert logits.shape == (tok.shape[0], tok.shape[1], int(g["V"]))
    ref = torch.from_numpy(g["logits"])
    assert rel_err(logits, ref) < 1e-3
    loss = m.compute_loss(logits, tgt)["ce_loss"]
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    am = logits.argmax(-1).cpu().numpy()
    top2 = ref.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).numpy()
    assert ((am == g["argmax"]) | (margin < 1e-4)).all()


def test_stage1_bf16_logits_vs_golden():
    g = golden("txl_small.npz")
    m = _model(g, torch.bfloat16).eval()
    tok = torch.from_numpy(g["tok"]).cuda()
    with torch.no_grad():
        logits, _ = m(tok, tuple())
    assert rms_rel(logits, torch.from_numpy(g["logits"])) < 2e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 8e-2)])
def test_stage1_gradients_vs_golden(dtype, tol):
    g = golden("txl_small.npz")
    m = _model(g, dtype).train()
    tok, tgt = torch.from_numpy(g["tok"]).cuda(), torch.from_numpy(g["tgt"]).cuda()
    m.zero_grad()
    logits, _ = m(tok, tuple())
    m.compute_loss(logits, tgt)["total_loss"].backward()
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
    assert n >= 6
    # fused step == autograd path
    g1 = m._flat_grad.clone()
    m.zero_grad()
    acc = m.train_step(tok, tgt)
    assert abs(float(acc[1] / acc[0]) - float(g["loss"])) < (1e-4 if dtype == torch.float32 else 3e-2)
    assert rel_err(m._flat_grad, g1) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_stage1_incremental_decode_with_memory_vs_golden():
    """generate(): primer of 5 tokens, then 20 single-token steps with mem_len 16 (hidden-state memory)."""
    g = golden("txl_small.npz")
    m = _model(g, torch.float32, mem_len=16).eval()
    tok = torch.from_numpy(g["tok"]).cuda()
    ref = torch.from_numpy(g["dec_logits"])
    mems = tuple()
    for step in range(21):
        inp = tok[:5, :1] if step == 0 else tok[4 + step:5 + step, :1]
        lg, mems = m.generate(inp, mems)
        assert lg.shape == (int(g["V"]),)
        assert len(mems) == int(g["L"]) + 1 and mems[0].shape[1:] == (1, 512) and mems[0].shape[0] <= 16
        assert rel_err(lg, ref[step]) < 1e-3, step
        assert int(lg.argmax()) == int(ref[step].argmax())           # greedy token identical


def test_stage1_decode_equals_full_forward_last_row():
    """property: with memory covering the whole prefix, step-wise decode == full forward (fp32)."""
    g = golden("txl_small.npz")
    m = _model(g, torch.float32, mem_len=64).eval()
    tok = torch.from_numpy(g["tok"]).cuda()[:, :1]
    with torch.no_grad():
        full, _ = m(tok, tuple())
    mems = tuple()
    for t in range(tok.shape[0]):
        lg, mems = m.generate(tok[t:t + 1], mems)
        assert rel_err(lg, full[t, 0]) < 1e-4, t


def test_stage1_dropout_consistent_fwd_bwd():
    """dropatt + renormalise (softmax over the kept keys): analytic gradient == finite differences of the SAME
    masked forward (kernel level, fp32)."""
    from emo_disentanger_b200 import ops
    B, H, T = 1, 8, 40
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, T, 3 * H * 64, generator=gen).to(DEV)
    r = torch.randn(T, H, 64, generator=gen).to(DEV)
    rw, rr = (0.3 * torch.randn(H, 64, generator=gen)).to(DEV), (0.3 * torch.randn(H, 64, generator=gen)).to(DEV)
    w = torch.randn(B, T, H * 64, generator=gen).to(DEV)
    d = H * 64

    def fwd(xx, r_):
        q, k, v = (xx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
        out = torch.empty(B, T, d, device=DEV)
        lse = torch.empty(B, H, T, device=DEV)
        ops.relattn_fwd(q, k, v, r_, rw, rr, out, lse, 0.125, 0.2, 11)
        return out, lse
    out, lse = fwd(x, r)
    dx = torch.empty_like(x)
    q, k, v = (x[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    dq, dk, dv = (dx[:, :, i * d:(i + 1) * d].unflatten(-1, (H, 64)) for i in range(3))
    dr, drw, drr = torch.zeros_like(r), torch.zeros_like(rw), torch.zeros_like(rr)
    ops.relattn_bwd(q, k, v, r, rw, rr, out, w, lse, dq, dk, dv, dr, drw, drr, 0.125, 0.2, 11)
    dirx = torch.randn(x.shape, generator=gen).to(DEV)
    dirr = torch.randn(r.shape, generator=gen).to(DEV)
    eps = 1e-2
    fp = (fwd(x + eps * dirx, r + eps * dirr)[0].double() * w.double()).sum()
    fm = (fwd(x - eps * dirx, r - eps * dirr)[0].double() * w.double()).sum()
    fd = float((fp - fm) / (2 * eps))
    an = float((dx.double() * dirx.double()).sum() + (dr.double() * dirr.double()).sum())
    assert abs(fd - an) / (abs(an) + 1e-6) < 2e-2


def test_stage1_training_with_dropout_learns_and_checkpoint_round_trip(tmp_path):
    from emo_disentanger_b200.optim import FusedAdam
    g = golden("txl_small.npz")
    tok, tgt = torch.from_numpy(g["tok"]).cuda(), torch.from_numpy(g["tgt"]).cuda()
    torch.manual_seed(3)
    m = _model(g, torch.bfloat16, dropout=0.1).train()
    opt = FusedAdam(m, lr=1e-3, max_grad_norm=0.5)
    losses = []
    for it in range(8):
        acc = m.train_step(tok, tgt)
        opt.step()
        losses.append(float(acc[1] / acc[0]))
    assert losses[-1] < losses[0] - 0.3
    path = tmp_path / "ep001_loss0.000_params.pt"
    torch.save(m.state_dict(), path)
    m2 = _model(g, torch.bfloat16)
    m2.load_state_dict(torch.load(path, map_location="cpu"))           # strict, as stage1 train.py:213-228
    assert torch.equal(m2._flat.cpu(), m._flat.cpu())
    assert len(m.state_dict()) == 6 + 11 * int(g["L"])


@pytest.mark.parametrize("B,Tq,Tk,p", [(4, 512, 512, 0.0), (4, 512, 512, 0.1), (12, 192, 192, 0.25), (16, 128, 320, 0.0)])
def test_relattn_tcgen05_equals_mma_sync(B, Tq, Tk, p):
    """the stage-1 relative-position attention through the tcgen05 kernels (position scores G = (q + r_r_bias) r^T by the
    tcgen05 GEMM, shifted by addressing in HBM, added to the score tiles inside the attention kernels) against the
    round-1 mma.sync kernels (band GEMM + skew in shared memory): out, lse, dq / dk / dv, dr and both bias gradients,
    with the same drop-and-renormalise mask"""
    from emo_disentanger_b200 import ops, _lib
    H, d = 8, 512
    g = torch.Generator().manual_seed(Tq + Tk)
    xq = (torch.randn(B, Tq, d, generator=g) * 0.8).to(DEV).to(torch.bfloat16)
    xkv = (torch.randn(B, Tk, 2 * d, generator=g) * 0.8).to(DEV).to(torch.bfloat16)
    r = (torch.randn(Tk, H, 64, generator=g) * 0.8).to(DEV).to(torch.bfloat16)
    rw, rr = (0.3 * torch.randn(H, 64, generator=g)).to(DEV), (0.3 * torch.randn(H, 64, generator=g)).to(DEV)
    dout = torch.randn(B, Tq, d, generator=g).to(DEV).to(torch.bfloat16)
    q = xq.unflatten(-1, (H, 64))
    k, v = xkv[:, :, :d].unflatten(-1, (H, 64)), xkv[:, :, d:].unflatten(-1, (H, 64))
    res = []
    for tc in (1, 0):
        _lib.lib().emo_attn_set_tc(tc)
        try:
            out = torch.empty(B, Tq, d, device=DEV, dtype=torch.bfloat16)
            lse = torch.empty(B, H, Tq, device=DEV)
            ops.relattn_fwd(q, k, v, r, rw, rr, out, lse, 0.125, p, 99)
            dxq, dxkv = torch.empty_like(xq), torch.empty_like(xkv)
            dq = dxq.unflatten(-1, (H, 64))
            dk, dv = dxkv[:, :, :d].unflatten(-1, (H, 64)), dxkv[:, :, d:].unflatten(-1, (H, 64))
            dr = torch.zeros(Tk, H, 64, device=DEV)
            drw, drr = torch.zeros(H, 64, device=DEV), torch.zeros(H, 64, device=DEV)
            if Tq == Tk:
                ops.relattn_bwd(q, k, v, r, rw, rr, out, dout, lse, dq, dk, dv, dr, drw, drr, 0.125, p, 99)
            torch.cuda.synchronize()
            res.append((out.float(), lse.clone(), dxq.float(), dxkv.float(), dr, drw, drr))
        finally:
            _lib.lib().emo_attn_set_tc(1)
    a, b = res
    assert torch.isfinite(a[0]).all()
    assert rms_rel(a[0], b[0]) < 1e-2, "out"
    fin = torch.isfinite(b[1])
    assert torch.equal(torch.isfinite(a[1]), fin) and float((a[1][fin] - b[1][fin]).abs().max()) < 3e-2, "lse"
    if Tq == Tk:
        for i, nm in ((2, "dq"), (3, "dkv"), (4, "dr"), (5, "d_r_w_bias"), (6, "d_r_r_bias")):
            assert torch.isfinite(a[i]).all(), nm
            assert rms_rel(a[i], b[i]) < 2.5e-2, (nm, rms_rel(a[i], b[i]))
