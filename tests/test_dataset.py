"""Stage-2 dataset sample assembly: oracle vs the golden outputs of the UNMODIFIED reference
`REMISkylineToMidiTransformerDataset` (tests/golden/make_dataset_golden.py), and the GPU token store vs both."""
import random
import numpy as np
import pytest
import torch

from helpers import golden
from oracle import dataset_oracle as DO

KEYS = ("dec_input", "dec_target", "track_mask", "chord_idx", "melody_idx")


def _pieces(g):
    return [([tuple(x) for x in g["p%d_mel" % p].tolist()], [tuple(x) for x in g["p%d_ch" % p].tolist()],
             g["p%d_tokens" % p].tolist()) for p in range(int(g["n_pieces"]))]


def test_oracle_matches_reference_golden():
    g = golden("dataset_small.npz")
    pieces, T = _pieces(g), int(g["seqlen"])
    assert len(g["picks"]) >= 60
    for i, (p, st, pk) in enumerate(g["picks"].tolist()):
        mel, ch, toks = pieces[p]
        out = DO.assemble(toks, mel, ch, st, T, int(g["pad"]), int(g["eos"]), g["is_chord"], g["is_note"], bool(pk))
        for k in KEYS:
            assert np.array_equal(out[k], g["s%d_%s" % (i, k)]), (i, p, st, pk, k)
        assert out["length"] == int(g["s%d_length" % i])


def test_oracle_edge_cases():
    # one bar, piece shorter than the window: everything after the piece is PAD, EOS closes the only bar
    mel, ch, toks = [(2, 5)], [(5, 9)], [10, 11, 1, 2, 3, 4, 5, 6, 7]
    out = DO.assemble(toks, mel, ch, 0, 16, 99, 50, np.zeros(100, int), np.zeros(100, int))
    assert out["dec_input"].tolist() == toks + [99] * 7 and out["length"] == 9
    assert out["dec_target"].tolist() == [99] * 5 + [5, 6, 7, 50] + [99] * 7
    assert out["track_mask"].tolist() == [0] * 5 + [1] * 4 + [0] * 7
    # admissible start bars: every bar from which at least half a window of events remains (contiguous from bar 0)
    assert DO.admissible_stbars(10, [(0, 1)], 16) == [0]
    assert DO.admissible_stbars(40, [(2, 5), (12, 15), (25, 28), (33, 36)], 16) == [0, 1, 2]


def _store(g, device, predict_key=False):
    from emo_disentanger_b200.data import Stage2TokenStore
    from emo_disentanger_b200.synth import synthetic_vocab
    e2i, _ = synthetic_vocab(int(g["V"]), 2)
    e2i = {k: v for k, v in e2i.items() if k != 'PAD_None'}
    i2e = {v: k for k, v in e2i.items()}
    pieces = [(mel, ch, [i2e[t] for t in toks]) for mel, ch, toks in _pieces(g)]
    return Stage2TokenStore(pieces, e2i, i2e, model_dec_seqlen=int(g["seqlen"]), predict_key=predict_key, device=device)


def test_token_store_tables_and_start_bars_cpu():
    g = golden("dataset_small.npz")
    st = _store(g, "cpu")
    pieces, T = _pieces(g), int(g["seqlen"])
    assert len(st) == len(pieces) and st.pad_token == int(g["pad"]) and st.eos_token == int(g["eos"])
    for p, (mel, ch, toks) in enumerate(pieces):
        a, b = int(st.piece_off[p]), int(st.piece_off[p + 1])
        assert st.tokens[a:b].tolist() == toks
        assert st.piece_admissible_stbars[p] == DO.admissible_stbars(len(toks), mel, T)
    assert np.array_equal(st.flags.numpy() & 1, g["is_chord"]) and np.array_equal((st.flags.numpy() >> 1) & 1, g["is_note"])
    with pytest.raises(Exception):
        st.batch([0])                                    # no CPU path: batches are assembled on the GPU


@pytest.mark.gpu
@pytest.mark.parametrize("predict_key", [False, True])
def test_token_store_batches_equal_reference_golden(predict_key):
    g = golden("dataset_small.npz")
    st = _store(g, "cuda", predict_key)
    picks = [(i, p, s) for i, (p, s, pk) in enumerate(g["picks"].tolist()) if bool(pk) == predict_key]
    batch = st.batch([p for _, p, _ in picks], [s for _, _, s in picks])          # one launch for all of them
    for row, (i, p, s) in enumerate(picks):
        for k in KEYS:
            assert np.array_equal(batch[k][row].cpu().numpy(), g["s%d_%s" % (i, k)]), (i, p, s, k)
        assert int(batch["length"][row]) == int(g["s%d_length" % i])
    assert batch["id"].tolist() == [p for _, p, _ in picks]


@pytest.mark.gpu
def test_token_store_full_size_vs_oracle_and_loader():
    """reference-sized window (T = 3072, emopia_finetune.yaml max_len), random pieces; the epoch iterator visits every
    piece once and draws start bars from the reference's admissible list with the `random` module"""
    from emo_disentanger_b200.data import Stage2TokenStore
    from emo_disentanger_b200.synth import synthetic_vocab
    V, T = 329, 3072
    e2i, _ = synthetic_vocab(V, 2)
    e2i = {k: v for k, v in e2i.items() if k != 'PAD_None'}
    i2e = {v: k for k, v in e2i.items()}
    rng = np.random.RandomState(3)
    pieces = [DO.synthetic_piece(e2i, nb, rng, lead_len=(5, 30), full_len=(20, 200)) for nb in (4, 20, 60, 90, 33)]
    st = Stage2TokenStore(pieces, e2i, i2e, model_dec_seqlen=T, device="cuda")
    is_chord, is_note = DO.vocab_flags(i2e, st.pad_token)
    seen = []
    random.seed(5)
    for batch in st.loader(batch_size=2, shuffle=True):
        for row, p in enumerate(batch["id"].tolist()):
            seen.append(p)
            mel, ch, ev = pieces[p]
            toks = [e2i[e] for e in ev]
            # recover the start bar the store drew: the first kept event after the 3-event header
            first = int(batch["dec_input"][row, mel[0][0] + 1]) if len(toks) > mel[0][0] + 1 else None
            cands = [b for b in st.piece_admissible_stbars[p]
                     if np.array_equal(DO.assemble(toks, mel, ch, b, T, st.pad_token, st.eos_token, is_chord, is_note)["dec_input"],
                                       batch["dec_input"][row].cpu().numpy())]
            assert cands, p
            ref = DO.assemble(toks, mel, ch, cands[0], T, st.pad_token, st.eos_token, is_chord, is_note)
            for k in KEYS:
                assert np.array_equal(batch[k][row].cpu().numpy(), ref[k]), (p, cands[0], k)
            assert int(batch["length"][row]) == ref["length"]
    assert sorted(seen) == list(range(len(pieces)))


# ------------------------------------------------------------------------------------------------------------
# stage-1 dataset: oracle vs the goldens of the UNMODIFIED SkylineFullSongTransformerDataset +
# collate_fn (tests/golden/make_stage1_dataset_golden.py), and the GPU token store vs both
# ------------------------------------------------------------------------------------------------------------
S1_KEYS = ("dec_inp", "dec_tgt", "inp_chord", "inp_melody")


def _s1_pieces(g):
    i2e = {i: str(n) for i, n in enumerate(g["vocab"].tolist())}
    e2i = {n: i for i, n in i2e.items()}
    pieces = [(g["p%d_bar_pos" % p].tolist(), [i2e[t] for t in g["p%d_tokens" % p].tolist()]) for p in range(int(g["n_pieces"]))]
    return e2i, i2e, pieces


def test_stage1_oracle_matches_reference_golden():
    g = golden("stage1_dataset.npz")
    e2i, i2e, pieces = _s1_pieces(g)
    pad = len(e2i)
    is_chord, is_note = DO.vocab_flags(i2e, pad)
    for ci, (seqlen, max_bars) in enumerate(g["configs"].tolist()):
        for p, (bp, ev) in enumerate(pieces):
            o = DO.stage1_assemble([e2i[e] for e in ev], bp, seqlen, max_bars, pad, e2i["EOS_None"], e2i["Bar_None"], is_chord, is_note)
            for k in S1_KEYS:
                assert np.array_equal(o[k], g["c%d_%s_0" % (ci, k)][p]), (ci, p, k)
            assert o["dec_seg_len"] == int(g["c%d_dec_seg_len_0" % ci][p])
    # the goldens cover: truncation (segment longer than the window), the max-bars cut (closing Bar instead of EOS)
    assert (g["c0_dec_seg_len_0"] > 64).any() and (g["c1_dec_tgt_0"] == e2i["Bar_None"]).any()


def test_stage1_oracle_edge_cases():
    bp, eos = DO.stage1_bar_positions([2, 6, 9], 12, 192)
    assert bp == [2, 6, 9, 11] and eos                                  # sentinel = position of EOS
    bp, eos = DO.stage1_bar_positions([2, 6, 9, 12], 12, 192)           # appended marker dropped
    assert bp == [2, 6, 9, 11]
    bp, eos = DO.stage1_bar_positions([2, 6, 10], 12, 192)              # trailing [Bar, EOS]: the empty bar goes
    assert bp == [2, 6, 9]
    bp, eos = DO.stage1_bar_positions([2, 6, 9], 12, 2)                 # more bars than max_bars: cut, closes with Bar
    assert bp == [2, 6, 9] and not eos
    assert DO.stage1_first_segment([2, 40, 80, 120], 64) == (0, 1)
    assert DO.stage1_first_segment([2, 90, 120, 150], 64) == (0, 1)     # a first bar longer than the window still ends at bar 1
    assert DO.stage1_first_segment([2, 10, 20], 64) == (0, 2)
    with pytest.raises(AssertionError):                                 # no event before the first bar: the reference asserts
        DO.stage1_assemble(list(range(8)), [0, 4], 64, 192, 99, 98, 97, np.zeros(100, int), np.zeros(100, int))


def test_stage1_token_store_tables_cpu():
    from emo_disentanger_b200.data import Stage1TokenStore
    g = golden("stage1_dataset.npz")
    e2i, i2e, pieces = _s1_pieces(g)
    for ci, (seqlen, max_bars) in enumerate(g["configs"].tolist()):
        st = Stage1TokenStore(pieces, e2i, i2e, model_dec_seqlen=seqlen, model_max_bars=max_bars, device="cpu")
        assert len(st) == len(pieces) and st.pad_token == len(e2i) and st.vocab_size == len(e2i) + 1
        assert st.seg_len.tolist() == g["c%d_dec_seg_len_0" % ci].tolist()
        for p, (bp, ev) in enumerate(pieces):
            rbp, eos = DO.stage1_bar_positions(bp, len(ev), max_bars)
            assert st.piece_bar_pos[p] == rbp and st.piece_segments[p] == [DO.stage1_first_segment(rbp, seqlen)]
            a, b = int(st.piece_off[p]), int(st.piece_off[p + 1])
            assert st.tokens[a:b].tolist() == DO.stage1_sample_tokens([e2i[e] for e in ev], rbp, eos, e2i["EOS_None"], e2i["Bar_None"])
        with pytest.raises(Exception):
            st.batch([0])                                               # no CPU path
    with pytest.raises(ValueError):
        Stage1TokenStore([([0, 4], ["Bar_None"] * 7 + ["EOS_None"])], e2i, i2e, device="cpu")


@pytest.mark.gpu
def test_stage1_token_store_batches_equal_reference_golden():
    from emo_disentanger_b200.data import Stage1TokenStore
    g = golden("stage1_dataset.npz")
    e2i, i2e, pieces = _s1_pieces(g)
    for ci, (seqlen, max_bars) in enumerate(g["configs"].tolist()):
        st = Stage1TokenStore(pieces, e2i, i2e, model_dec_seqlen=seqlen, model_max_bars=max_bars, device="cuda")
        order = list(range(len(pieces)))[::-1]
        batch = st.batch(order)                                         # one launch for all pieces
        for row, p in enumerate(order):
            for k in S1_KEYS:
                assert np.array_equal(batch[k + "_0"][row].cpu().numpy(), g["c%d_%s_0" % (ci, k)][p]), (ci, p, k)
        assert batch["dec_seg_len_0"].tolist() == g["c%d_dec_seg_len_0" % ci][order].tolist()
        assert batch["id"].tolist() == order and int(max(batch["n_seg"])) == 1
        assert st.batch([])["dec_inp_0"].shape == (0, seqlen)           # empty batch
    # reference-sized window through the epoch iterator: every piece once, rows equal to the oracle
    rng = np.random.RandomState(2)
    big = [DO.synthetic_stage1_piece(rng, nb, bar_len=(10, 40)) for nb in (3, 40, 150, 260, 17)]
    from emo_disentanger_b200.data import formats as F
    e2, i2 = F.build_dictionary([ev for _, ev in big], relative=True, **F.VOCAB_FLAGS["stage1_lead_sheet"])
    st = Stage1TokenStore(big, e2, i2, model_dec_seqlen=2400, device="cuda")
    is_chord, is_note = DO.vocab_flags(i2, st.pad_token)
    seen = []
    random.seed(1)
    for batch in st.loader(batch_size=2):
        for row, p in enumerate(batch["id"].tolist()):
            seen.append(p)
            bp, ev = big[p]
            o = DO.stage1_assemble([e2[F.event_name(e)] for e in ev], bp, 2400, 192, st.pad_token, st.eos_token, st.bar_token, is_chord, is_note)
            for k in S1_KEYS:
                assert np.array_equal(batch[k + "_0"][row].cpu().numpy(), o[k]), (p, k)
            assert int(batch["dec_seg_len_0"][row]) == o["dec_seg_len"]
    assert sorted(seen) == list(range(len(big)))


@pytest.mark.parametrize("stage", [1, 2])
def test_token_store_loader_shards_an_epoch_across_ranks(stage, monkeypatch):
    """data parallel: rank r of `world` takes batches r::world of ONE shuffled order drawn from
    Random(seed + epoch) -- never the process-global `random` state, which differs per rank -- disjoint, the same number
    of batches on every rank.  Host logic only (batch() is stubbed)."""
    from emo_disentanger_b200.data import Stage1TokenStore, Stage2TokenStore
    if stage == 1:
        g = golden("stage1_dataset.npz")
        e2i, i2e, pieces = _s1_pieces(g)
        st = Stage1TokenStore(pieces, e2i, i2e, model_dec_seqlen=64, device="cpu")
    else:
        st = _store(golden("dataset_small.npz"), "cpu")
    monkeypatch.setattr(type(st), "batch", lambda self, idx, *a: list(idx))
    per_rank = []
    for r in range(2):
        random.seed(1000 + r)                  # ranks do NOT share the global RNG state
        per_rank.append(list(st.loader(batch_size=2, shuffle=True, rank=r, world=2, seed=9, epoch=1)))
    whole = list(st.loader(batch_size=2, shuffle=True, seed=9, epoch=1))
    n_even = len(whole) // 2 * 2
    assert whole[0:n_even:2] == per_rank[0] and whole[1:n_even:2] == per_rank[1]
    assert len(per_rank[0]) == len(per_rank[1])
    flat = [p for b in whole for p in b]
    assert sorted(flat) == list(range(len(st))) and all(len(b) <= 2 for b in whole)
    random.seed(123)
    dropped = list(st.loader(batch_size=4, shuffle=True, drop_last=True))
    assert all(len(b) == 4 for b in dropped) and len(dropped) == len(st) // 4
