"""Event -> MIDI post-processing: the score (notes, tempo changes, chord / bar markers, block chords)
against the UNMODIFIED reference converter run over stand-in containers (tests/golden/make_midi_golden.py), the
functional -> absolute conversion against `convert_key.degree2pitch`, and a write/read round trip of the MIDI file.
The file BYTES are not pinned against miditoolkit (not installed).  CPU only."""
import json
import os
import random

import pytest

from emo_disentanger_b200.data import midi_out as M

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "midi_score.json")))


def _norm(score):
    return {"instruments": [[list(n) for n in ins] for ins in score["instruments"]],
            "tempos": [list(t) for t in score["tempos"]], "markers": [list(m) for m in score["markers"]],
            "max_tick": score["max_tick"]}


@pytest.mark.parametrize("case", GOLD["score"], ids=lambda c: "s%d-%s-%s" % (c["stage"], c["mode"], c["key"]))
def test_events_to_score_matches_reference(case):
    enforce = [tuple(t) for t in case["enforce_tempos"]] if "enforce_tempos" in case else None
    got = M.events_to_score(case["key"], case["events"], mode=case["mode"], play_chords=case["play_chords"],
                            enforce_tempos=enforce)
    assert _norm(got) == case["result"]


@pytest.mark.parametrize("case", GOLD["degree"], ids=lambda c: "seed%d" % c["seed"])
def test_degree_to_pitch_and_minor_table_match_reference(case):
    table = M.minor_degree_table(random.Random(case["seed"]))       # the reference draws at import of convert_key
    assert table == case["minor_table"]
    for k, want in case["pitch"].items():
        key, octave, roman = k.split("|")
        assert M.degree_to_pitch(key, int(octave), roman, table) == want, k
    with pytest.raises(NameError):
        M.degree_to_pitch("H", 4, "I")


def test_relative_to_absolute_and_bar_split():
    table = dict(M.MINOR_DEGREE)
    ev = ["Track_LeadSheet", "Bar_None", "Beat_0", "Chord_V_7", "Note_Octave_4", "Note_Degree_I", "Note_Duration_480",
          "Track_Full", "Beat_0", "Chord_Conti_Conti", "Note_Octave_9", "Note_Degree_VII", "Note_Duration_120", "Note_Velocity_60",
          "Track_LeadSheet", "Bar_None", "Beat_8", "Chord_None_None", "Note_Octave_0", "Note_Degree_I", "Note_Duration_240",
          "Track_Full", "Beat_8", "Chord_III_m", "Note_Octave_5", "Note_Degree_III", "Note_Duration_240", "Note_Velocity_70", "EOS_None"]
    major = M.relative_to_absolute("Key_D", ev)
    assert major[3] == "Chord_7_7" and major[4] == "Note_Pitch_%d" % (4 * 12 + 2 + 0)
    assert "Chord_Conti_Conti" in major and "Chord_None_None" in major
    assert "Note_Pitch_108" in major and "Note_Pitch_21" in major           # clamped to the piano range
    assert not any("Note_Octave" in e or "Note_Degree" in e for e in major) and len(major) == len(ev) - 4
    minor = M.relative_to_absolute("Key_a", ev, minor_degree=table)
    assert "Chord_%d_m" % table["III"] in minor and minor[4] == "Note_Pitch_%d" % (4 * 12 + 9)
    with pytest.raises(KeyError):
        M.relative_to_absolute("Key_C", ["Chord_Conti_Conti"], keep_conti_chords=False)     # the stage-1 variant
    bars = M.full_track_bars(major)
    assert len(bars) == 2 and bars[0][0] == "Beat_0" and bars[1][-1] == "EOS_None" and "Track_Full" not in sum(bars, [])


@pytest.mark.parametrize("case", [GOLD["score"][0], GOLD["score"][3]], ids=["full", "lead+chords"])
def test_midi_file_round_trip(case, tmp_path):
    score = M.events_to_score(case["key"], case["events"], mode=case["mode"], play_chords=case["play_chords"])
    p = tmp_path / "x.mid"
    M.write_midi(p, score)
    raw = open(p, "rb").read()
    assert raw[:4] == b"MThd" and raw.count(b"MTrk") == 1 + len(score["instruments"])
    back = M.read_midi(p)
    assert back["ticks_per_beat"] == 480 and back["format"] == 1
    assert back["tempos"] == [tuple(t) for t in score["tempos"]]
    assert sorted(back["markers"]) == sorted(tuple(m) for m in score["markers"])
    assert len(back["instruments"]) == len([i for i in score["instruments"] if i])
    for got, want in zip(back["instruments"], [i for i in score["instruments"] if i]):
        assert sorted(got) == sorted(tuple(n) for n in want)
    assert back["max_tick"] == score["max_tick"]


def test_score_without_notes_is_an_error_like_the_reference():
    with pytest.raises(ValueError):
        M.events_to_score("Key_C", ["Bar_None", "Beat_0", "Chord_0_M"], mode="full")


def test_inference_scripts_write_the_reference_outputs(tmp_path):
    """the per-piece files of stage-1 inference.py:252-283 and the accompaniment .mid of stage-2 inference.py:462-479"""
    from emo_disentanger_b200.scripts.stage1_inference import write_lead_sheet_outputs
    from emo_disentanger_b200.scripts.stage2_inference import write_accompaniment_midi
    lead = ["Emotion_Positive", "Key_G", "Bar_None", "Beat_0", "Chord_I_M", "Note_Octave_5", "Note_Degree_I", "Note_Duration_480",
            "Beat_8", "Chord_V_7", "Note_Octave_5", "Note_Degree_II", "Note_Duration_960", "Bar_None", "Beat_0", "Chord_I_M",
            "Note_Octave_4", "Note_Degree_VII", "Note_Duration_1920", "EOS_None"]
    assert write_lead_sheet_outputs(str(tmp_path), "samp_00_Positive", lead, "functional", "lead_sheet") == \
        ["samp_00_Positive_roman.txt", "samp_00_Positive.txt", "samp_00_Positive.mid"]
    assert open(tmp_path / "samp_00_Positive_roman.txt").read().splitlines() == lead[1:]
    absolute = open(tmp_path / "samp_00_Positive.txt").read().splitlines()
    assert absolute[0] == "Key_G" and "Note_Pitch_%d" % (5 * 12 + 7) in absolute and "Chord_7_7" in absolute
    mid = M.read_midi(tmp_path / "samp_00_Positive.mid")
    assert mid["tempos"] == [(110, 0)] and len(mid["instruments"]) == 2                  # melody + block chords
    assert sorted(n[1] for n in mid["instruments"][0]) == sorted([67, 69, 48 + 7 + 11])
    assert ("Chord-G_M", 0) in mid["markers"] and ("Chord-D_7", 960) in mid["markers"] and ("Bar-1", 0) in mid["markers"]
    # REMI representation: no roman file; a sequence without a complete note leaves the text files only
    assert write_lead_sheet_outputs(str(tmp_path), "samp_01_Negative", ["Emotion_Negative", "Bar_None", "Beat_0", "Chord_0_M"],
                                    "remi", "lead_sheet") == ["samp_01_Negative.txt"]
    full = ["Emotion_Q1", "Key_G", "Tempo_110", "Track_LeadSheet", "Bar_None", "Beat_0", "Note_Octave_5", "Note_Degree_I",
            "Note_Duration_480", "Track_Full", "Bar_None", "Beat_0", "Tempo_110", "Chord_I_M", "Note_Octave_3", "Note_Degree_I", "Note_Duration_960",
            "Note_Velocity_60", "Note_Octave_4", "Note_Degree_V", "Note_Duration_480", "Note_Velocity_72", "Track_LeadSheet",
            "Bar_None", "Beat_0", "Note_Octave_5", "Note_Degree_III", "Note_Duration_480", "Track_Full", "Bar_None", "Beat_4", "Note_Octave_3",
            "Note_Degree_IV", "Note_Duration_240", "Note_Velocity_50", "EOS_None"]
    assert write_accompaniment_midi(str(tmp_path), "samp_00_Q1_full", "Key_G", full, "functional")
    acc = M.read_midi(tmp_path / "samp_00_Q1_full.mid")
    # only the Full-track notes (each span opens with its own Bar event, midi2events_emopia.py:492-493)
    assert sorted((n[1], n[0], n[2]) for n in acc["instruments"][0]) == sorted([(43, 60, 0), (62, 72, 0), (48, 50, 1920 + 480)])
    assert not write_accompaniment_midi(str(tmp_path), "none", "Key_C", ["Track_LeadSheet", "Bar_None", "Track_Full", "Bar_None", "Beat_0"], "remi")
    # a note before any Bar event would sit at a negative tick: refused, not written
    assert not write_accompaniment_midi(str(tmp_path), "neg", "Key_C", ["Track_Full", "Beat_0", "Note_Pitch_60", "Note_Duration_120", "Note_Velocity_60"], "remi")
