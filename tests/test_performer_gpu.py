"""GPU parity of the stage-2 Performer module (through the reference-facing nn.Module API) against
the oracle's committed golden vectors and against the oracle run on the same seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import golden, rel_err, rms_rel, load_seeded
from oracle import performer_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(g, dtype, dropout=0.0):
    from emo_disentanger_b200.stage2 import MusicPerformer
    V, L = int(g["V"]), int(g["L"])
    m = MusicPerformer(V, L, 8, 512, 2048, 512, dropout=dropout, use_segment_emb=True, n_segment_types=2,
                       favor_feature_dims=128, compute_dtype=dtype)
    load_seeded(m, PO.performer_state_shapes(V, L), int(g["seed"]))
    m = m.cuda()
    m.fixed_omegas = torch.from_numpy(g["omegas"]).cuda()
    return m


def _inputs(g):
    return (torch.from_numpy(g["tok"]).cuda(), torch.from_numpy(g["seg"]).cuda(), torch.from_numpy(g["tgt"]).cuda())


def test_fp32_logits_loss_argmax_vs_golden():
    g = golden("performer_small.npz")
    m = _model(g, torch.float32).eval()
    tok, seg, tgt = _inputs(g)
    with torch.no_grad():
        logits = m(tok, seg_inp=seg)
    ref = torch.from_numpy(g["logits"])
    assert rel_err(logits, ref) < 1e-3                      # north-star tolerance: 1e-3 rel on fp32 logits
    loss = m.compute_loss(logits, tgt)["recons_loss"]
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    am = logits.argmax(-1).cpu().numpy()
    top2 = ref.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).numpy()
    assert ((am == g["argmax"]) | (margin < 1e-4)).all()    # greedy tokens identical (ties excepted)
    last = m(tok, seg_inp=seg, keep_last_only=True)
    assert torch.equal(last, logits[:, -1, :])


def test_bf16_hidden_and_logits_vs_golden():
    g = golden("performer_small.npz")
    m = _model(g, torch.bfloat16).eval()
    tok, seg, tgt = _inputs(g)
    with torch.no_grad():
        hid, _ = m._forward_hidden(tok, seg, save=False)
        logits = m(tok, seg_inp=seg)
    ref_h = torch.from_numpy(g["hidden_last"]).view(-1, 512)
    assert rms_rel(hid.float(), ref_h) < 1e-2                # north-star tolerance: 1e-2 on bf16 hidden states
    assert rms_rel(logits, torch.from_numpy(g["logits"])) < 2e-2
    loss = m.compute_loss(logits, tgt)["recons_loss"]
    assert abs(float(loss) - float(g["loss"])) < 2e-2


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 6e-2)])      # bf16 measures <= 3e-2 except one bias gradient (linear1.bias of layer 1, heavy cancellation: 5.1e-2)
def test_gradients_vs_golden(dtype, tol):
    g = golden("performer_small.npz")
    m = _model(g, dtype).train()          # dropout p = 0 -> deterministic
    tok, seg, tgt = _inputs(g)
    m.zero_grad()
    logits = m(tok, seg_inp=seg)
    losses = m.compute_loss(logits, tgt)
    losses["total_loss"].backward()
    named = dict(m.named_parameters())
    for key in g.files:
        if key.startswith("grad:"):
            name = key[5:]
            e = rms_rel(named[name].grad, torch.from_numpy(g[key]))
            assert e < tol, "%s rms rel err %.3e" % (name, e)
        elif key.startswith("gradslice:"):
            name = key[10:]
            e = rms_rel(named[name].grad.reshape(-1)[:2048], torch.from_numpy(g[key]))
            assert e < tol, "%s rms rel err %.3e" % (name, e)
        elif key.startswith("gradnorm:"):
            name = key[9:]
            assert abs(float(named[name].grad.norm()) / float(g[key]) - 1) < tol, name


def test_train_step_equals_autograd_path():
    g = golden("performer_small.npz")
    m = _model(g, torch.float32).train()
    tok, seg, tgt = _inputs(g)
    m.zero_grad()
    m.compute_loss(m(tok, seg_inp=seg), tgt)["total_loss"].backward()
    g1 = m._flat_grad.clone()
    m.zero_grad()
    acc = m.train_step(tok, seg, tgt)
    assert abs(float(acc[1] / acc[0]) - float(g["loss"])) < 1e-4
    assert rel_err(m._flat_grad, g1) < 1e-5
    # gradient accumulation: a second step doubles the buffer
    m.train_step(tok, seg, tgt)
    assert rel_err(m._flat_grad, 2 * g1) < 1e-5


def test_oracle_same_inputs_other_shape():
    """fresh seeded inputs at a ragged length (T not a multiple of the chunk), fp32 mode."""
    from emo_disentanger_b200.stage2 import MusicPerformer
    V, L, B, T = 216, 1, 3, 77
    m = MusicPerformer(V, L, 8, 512, 2048, 512, use_segment_emb=False, favor_feature_dims=128,
                       compute_dtype=torch.float32)
    sd = load_seeded(m, {k: v for k, v in PO.performer_state_shapes(V, L).items() if "segemb" not in k}, 21)
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(1)
    om = torch.randn(L, 64, 64, generator=gen)
    m.fixed_omegas = om.cuda()
    tok = torch.randint(0, V - 1, (B, T), generator=gen)
    with torch.no_grad():
        out = m(tok.cuda())
    ref = PO.performer_forward(sd, tok, None, [om[0]], L, 8, 512)
    assert rel_err(out, ref) < 1e-3


def test_training_with_dropout_learns_and_is_seed_deterministic():
    from emo_disentanger_b200.optim import FusedAdam
    g = golden("performer_small.npz")
    tok, seg, tgt = _inputs(g)
    losses = []
    for rep in range(2):
        torch.manual_seed(1234)
        m = _model(g, torch.bfloat16, dropout=0.1).train()
        opt = FusedAdam(m, lr=1e-3, max_grad_norm=0.5)
        cur = []
        for it in range(8):
            acc = m.train_step(tok, seg, tgt)
            opt.step()
            cur.append(float(acc[1] / acc[0]))
        losses.append(cur)
    # same seed -> same dropout masks -> same trajectory, up to the summation order of the fp32 atomics
    # in the split-K weight-gradient / bias-gradient reductions
    assert max(abs(a - b) for a, b in zip(*losses)) < 2e-2
    assert losses[0][-1] < losses[0][0] - 0.3           # it learns
    assert float(m._flat_grad.abs().max()) == 0         # fused step zeroed the gradient buffer


def test_reference_checkpoint_round_trip(tmp_path):
    g = golden("performer_small.npz")
    m = _model(g, torch.bfloat16)
    path = tmp_path / "ep001_loss0.000_params.pt"
    torch.save(m.state_dict(), path)
    from emo_disentanger_b200.stage2 import MusicPerformer
    m2 = MusicPerformer(int(g["V"]), int(g["L"]), 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2,
                        favor_feature_dims=128).cuda()
    pre = torch.load(path, map_location="cpu")
    pre = {k: v for k, v in pre.items() if "feature_map.omega" not in k}       # as train.py:306-308
    sd = m2.state_dict(); sd.update(pre); m2.load_state_dict(sd)
    assert torch.equal(m2._flat, m._flat)
