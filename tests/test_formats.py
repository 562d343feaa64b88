"""On-disk formats: the closed-form vocabulary and the dictionary order against the reference's
`events2words` (goldens minted by tests/golden/make_formats_golden.py), pickle / lead-sheet round trips, and the
token store fed from files written in the reference layout.  CPU only."""
import json
import os
import pickle

import pytest

from emo_disentanger_b200.data import formats as F

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "formats.json")))


@pytest.mark.parametrize("case", GOLD["vocab"], ids=lambda c: "%s-v%d-e%d-t%d-n%d" % (
    "functional" if c["relative"] else "remi", c["add_velocity"], c["add_emotion"], c["add_tempo"], c["num_emotion"]))
def test_full_vocab_matches_reference(case):
    got = F.full_vocab(add_velocity=case["add_velocity"], add_emotion=case["add_emotion"], add_tempo=case["add_tempo"],
                       num_emotion=case["num_emotion"], relative=case["relative"])
    assert got == case["events"]                  # same events in the same emission order
    assert len(set(got)) == len(got)


def test_build_dictionary_matches_reference_order():
    d = GOLD["dictionary"]
    e2i, i2e = F.build_dictionary(d["observed"], **d["flags"])
    assert e2i == d["event2idx"]
    assert {str(k): v for k, v in i2e.items()} == d["idx2event"]
    # string-typed events give the same dictionary as dict-typed ones
    as_str = [[F.event_name(e) for e in seq] for seq in d["observed"]]
    assert F.build_dictionary(as_str, **d["flags"])[0] == e2i


def test_dictionary_and_piece_round_trip(tmp_path):
    d = GOLD["dictionary"]
    e2i, i2e = F.build_dictionary(d["observed"], **d["flags"])
    p = tmp_path / "dictionary.pkl"
    F.save_dictionary(p, e2i, i2e)
    assert pickle.load(open(p, "rb")) == (e2i, i2e)              # the reference's tuple layout
    e2, i2, V = F.load_dictionary(p)
    assert (e2, i2) == (e2i, i2e) and V == len(e2i) + 1          # + PAD
    F.save_piece(tmp_path / "s2.pkl", [[0, 3]], [[0, 5]], d["observed"][0])
    F.save_piece(tmp_path / "s1.pkl", [0, 7], d["observed"][1])
    s2, s1 = F.load_piece(tmp_path / "s2.pkl"), F.load_piece(tmp_path / "s1.pkl")
    assert set(s2) == {"lead_pos", "full_pos", "events"} and s2["events"] == d["observed"][0]
    assert set(s1) == {"bar_pos", "events"} and s1["bar_pos"] == [0, 7]
    assert F.piece_files(str(tmp_path)) == [str(tmp_path / n) for n in ("dictionary.pkl", "s1.pkl", "s2.pkl")]
    assert F.piece_files(str(tmp_path), ["s2.pkl"]) == [str(tmp_path / "s2.pkl")]


def test_lead_sheet_text_round_trip(tmp_path):
    e2i = {n: i for i, n in enumerate(["Bar_None", "Beat_0", "Beat_8", "Chord_I_M", "Key_G", "Note_Degree_V", "Emotion_Positive"])}
    events = ["Key_G", "Emotion_Positive", "Bar_None", "Beat_0", "Chord_I_M", "Bar_None", "Beat_8", "Note_Degree_V", "Bar_None"]
    p = tmp_path / "samp_00_Positive_roman.txt"
    F.write_events(p, events)
    key, bars = F.read_lead_sheet(p, e2i)
    assert key == "Key_G"
    assert bars == [[0, 1, 3], [0, 2, 5], [0]]                   # header before the first bar dropped; empty last bar kept
    F.write_events(tmp_path / "samp_01_Negative.txt", events[2:])
    key2, bars2 = F.read_lead_sheet(tmp_path / "samp_01_Negative.txt", e2i)
    assert key2 == "Key_C" and bars2 == bars                     # no key line -> C
    with pytest.raises(KeyError):
        F.read_lead_sheet(p, {"Bar_None": 0})                    # out-of-vocabulary events are an error, as in the reference
    F.write_events(tmp_path / "samp_00_Q1_full.txt", events)
    assert F.lead_sheet_files(str(tmp_path), "functional") == [str(p)]
    assert F.lead_sheet_files(str(tmp_path), "remi") == [str(tmp_path / "samp_00_Positive_roman.txt"), str(tmp_path / "samp_01_Negative.txt")]
    assert F.emotions_for("samp_00_Positive_roman.txt") == ["Q1", "Q4"]
    assert F.emotions_for("samp_01_Negative.txt") == ["Q2", "Q3"]
    assert F.emotions_for("x_Q3.txt") == ["Q3"]
    with pytest.raises(ValueError):
        F.emotions_for("nothing.txt")


def test_token_store_tables_from_reference_layout_files(tmp_path):
    """files written in the reference layout -> Stage2TokenStore.from_files (host tables only; no GPU needed)"""
    from emo_disentanger_b200.data.token_store import Stage2TokenStore
    seqs = [[{"name": "Bar", "value": None}, {"name": "Track", "value": "LeadSheet"}, {"name": "Note_Degree", "value": "I"},
             {"name": "Track", "value": "Full"}, {"name": "Chord", "value": "I_M"}, {"name": "EOS", "value": None}]]
    e2i, i2e = F.build_dictionary(seqs, relative=True, **F.VOCAB_FLAGS["stage2_full_song"])
    F.save_dictionary(tmp_path / "dictionary.pkl", e2i, i2e)
    os.makedirs(tmp_path / "events")
    F.save_piece(tmp_path / "events" / "a.pkl", [[0, 1]], [[3, 5]], seqs[0])
    st = Stage2TokenStore.from_files(F.piece_files(str(tmp_path / "events")), str(tmp_path / "dictionary.pkl"),
                                     model_dec_seqlen=16, device="cpu")
    assert st.vocab_size == len(e2i) + 1 and st.pad_token == len(e2i)
    assert st.tokens.tolist() == [e2i[F.event_name(e)] for e in seqs[0]]
    assert st.piece_ids == ["a"] and st.piece_admissible_stbars == [[0]]
