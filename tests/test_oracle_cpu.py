"""CPU suite: oracle vs committed goldens, oracle self-consistency, ABI surface, host logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import performer_oracle as PO, gpt2_oracle as GO, txl_oracle as TO, sampling_oracle as SO
from helpers import golden, rel_err, wsum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_performer_oracle_matches_golden():
    g = golden("performer_small.npz")
    V, L = int(g["V"]), int(g["L"])
    shapes = PO.performer_state_shapes(V, L)
    sd = PO.seeded_state(shapes, int(g["seed"]))
    assert abs(wsum(sd, shapes) - float(g["wsum"])) < 1e-6 * float(g["wsum"]), "seeded weights drifted"
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    omegas = [torch.from_numpy(o) for o in g["omegas"]]
    logits = PO.performer_forward(sd, torch.from_numpy(g["tok"]), torch.from_numpy(g["seg"]), omegas, L, 8, 512)
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-5
    loss = PO.ce_loss(logits, torch.from_numpy(g["tgt"]), V - 1)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    assert (logits.argmax(-1).numpy() == g["argmax"]).all()


def test_gpt2_oracle_matches_golden():
    g = golden("gpt2_small.npz")
    V, L = int(g["V"]), int(g["L"])
    shapes = GO.gpt2_state_shapes(V, L)
    sd = PO.seeded_state(shapes, int(g["seed"]))
    sd["pe.pe"] = PO.sinusoid_pe(12000, 512)
    logits = GO.gpt2_forward(sd, torch.from_numpy(g["tok"]), torch.from_numpy(g["seg"]), L, 8, 512)
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-5


def test_txl_oracle_matches_golden():
    g = golden("txl_small.npz")
    V, L = int(g["V"]), int(g["L"])
    sd = PO.seeded_state(TO.txl_state_shapes(V, L), int(g["seed"]))
    tok = torch.from_numpy(g["tok"])
    logits, _ = TO.txl_forward(sd, tok, None, L, 8, 512, 0)
    assert rel_err(logits, torch.from_numpy(g["logits"])) < 1e-5
    mems = None
    for step in range(21):
        inp = tok[:5, :1] if step == 0 else tok[4 + step:5 + step, :1]
        lg, mems = TO.txl_generate(sd, inp, mems, L, 8, 512, 16)
        assert rel_err(lg, torch.from_numpy(g["dec_logits"][step])) < 1e-5
        assert mems[0].shape[0] == min(16, 5 + step)


def test_txl_decode_equals_full_forward_within_memory():
    """incremental decode with a large enough memory == full-sequence forward (causality + rel-pos)."""
    V, L = 50, 2
    sd = PO.seeded_state(TO.txl_state_shapes(V, L), 3)
    tok = torch.randint(0, V - 1, (12, 1))
    full, _ = TO.txl_forward(sd, tok, None, L, 8, 512, 0)
    mems = None
    for t in range(12):
        lg, mems = TO.txl_generate(sd, tok[t:t + 1], mems, L, 8, 512, 64)
        assert rel_err(lg, full[t, 0]) < 1e-4


def test_sampling_oracle_matches_reference_golden():
    g = golden("sampling_ref.npz")
    for i in range(len(g["V"])):
        V = int(g["V"][i])
        logits = g["logits"][i][:V].astype(np.float32)
        try:
            w = SO.sample(logits, float(g["t"][i]), float(g["p"][i]), float(g["u"][i]))
        except IndexError:
            w = -1
        assert w == int(g["word"][i])


def test_sampling_edge_cases():
    # one dominant token: exactly one index above p -> the reference raises IndexError (kept)
    logits = np.full(8, -50.0, dtype=np.float32)
    logits[3] = 50.0
    cand, cp = None, None
    with pytest.raises(IndexError):
        SO.nucleus_candidates(SO.temperature_probs(logits[:1], 1.0), 0.9)
    # uniform distribution: nothing exceeds p=1.0 -> top-3 fallback
    cand, cp = SO.nucleus_candidates(np.full(10, 0.1, dtype=np.float32), 1.5)
    assert len(cand) == 3 and abs(cp.sum() - 1) < 1e-12
    assert SO.greedy(np.array([1.0, 3.0, 3.0, 2.0])) == 1


def test_causal_product_forms_agree():
    q, k, v = (torch.randn(2, 37, 2, 64, dtype=torch.float64) for _ in range(3))
    om = PO.draw_omega(64, 64, dtype=torch.float64)
    a, _ = PO.causal_linear_attention(q, k, v, om, sequential=True)
    b, _ = PO.causal_linear_attention(q, k, v, om, sequential=False)
    assert rel_err(a, b) < 1e-12
    # prefix property: output at position l does not depend on later tokens
    c, _ = PO.causal_linear_attention(q[:, :20], k[:, :20], v[:, :20], om)
    assert rel_err(c, a[:, :20]) < 1e-12


def test_favor_converges_to_softmax():
    torch.manual_seed(0)
    q, k = torch.randn(1, 16, 1, 64, dtype=torch.float64) * 0.5, torch.randn(1, 16, 1, 64, dtype=torch.float64) * 0.5
    exact = torch.exp(torch.einsum("nlhe,njhe->nlj", q, k) / 8.0)
    om = PO.draw_omega(64, 16384, dtype=torch.float64)
    Q, K = PO.favor_features(q, om, 32768), PO.favor_features(k, om, 32768)
    approx = torch.einsum("nlhi,njhi->nlj", Q, K)
    assert float(((approx - exact).abs() / exact).mean()) < 0.05


def test_orthogonal_omega_block():
    om = PO.draw_omega(64, 64, orthogonal=True, dtype=torch.float64)
    g = om.T @ om
    off = g - torch.diag(torch.diag(g))
    assert off.abs().max() < 1e-9          # columns orthogonal, scaled by chi-distributed norms


# ---- C ABI surface -----------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "emo_b200.h")).read()
    declared = set(re.findall(r"\b(emo_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"emo_status", "emo_dtype", "emo_gemm_op", "emo_act", "emo_epilogue"}
    assert len(declared) >= 20
    so = os.path.join(ROOT, "emo_disentanger_b200", "libemo_b200.so")
    if not os.path.exists(so):
        from emo_disentanger_b200 import build
        build.build()
    lib = ctypes.CDLL(so)
    for name in sorted(declared):
        assert hasattr(lib, name), "missing export: " + name
    from emo_disentanger_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert _lib.lib().emo_version() >= 100


def test_product_path_refuses_cpu_tensors():
    from emo_disentanger_b200 import ops, _lib
    x = torch.zeros(4, 512)
    with pytest.raises(_lib.EmoError):
        ops.linear_fwd(x, torch.zeros(8, 512), torch.zeros(4, 8))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "emo_disentanger_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src, "%s references the oracle" % f


# ---- host logic ---------------------------------------------------------------------------------
def test_flat_module_state_dict_and_views():
    from emo_disentanger_b200.stage2 import MusicPerformer
    m = MusicPerformer(329, 2, 8, 512, 2048, 512, use_segment_emb=True, n_segment_types=2, favor_feature_dims=128)
    shapes = PO.performer_state_shapes(329, 2)
    sd = m.state_dict()
    for k, shp in shapes.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert tuple(sd["pe.pe"].shape) == (12000, 1, 512)
    assert "transformer_decoder.decoder_layers.1.attention.inner_attention.feature_map.omega" in sd
    # parameters are views of one flat buffer; q,k,v are adjacent (packed QKV GEMM)
    base = m._flat.data_ptr()
    q = m.transformer_decoder.decoder_layers._modules["0"].attention.query_projection.weight
    k = m.transformer_decoder.decoder_layers._modules["0"].attention.key_projection.weight
    assert k.data_ptr() - q.data_ptr() == 512 * 512 * 4 and q.data_ptr() >= base
    # load_state_dict writes through to the flat buffer
    new = PO.seeded_state(shapes, 5)
    msd = m.state_dict(); msd.update(new); m.load_state_dict(msd)
    assert torch.equal(m._qkv_w(m._flat, 0)[512:1024], new["transformer_decoder.decoder_layers.0.attention.key_projection.weight"])
    # gradients are views of the flat gradient buffer and survive zero_grad
    m.zero_grad()
    assert q.grad.data_ptr() == m._flat_grad.data_ptr() + (q.data_ptr() - base)
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == sum(int(np.prod(s)) for s in shapes.values())


def test_warmup_cosine_matches_torch_scheduler():
    from emo_disentanger_b200.optim import WarmupCosine

    class O:  # minimal optimizer stand-in
        param_groups = [{"lr": 0.0}]
    w = WarmupCosine(O, 1e-5, 1e-6, 200, 500000)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-5)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, 500000, eta_min=1e-6)
    for step in (1, 100, 199):
        assert abs(w.update(step) - 1e-5 * step / 200) < 1e-18
    import warnings
    for step in (200, 1000, 250000):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sched.step(step - 200)
        assert abs(w.update(step) - opt.param_groups[0]["lr"]) < 1e-12
