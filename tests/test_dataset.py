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
