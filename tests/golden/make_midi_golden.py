"""Mint tests/golden/midi_score.json: the UNMODIFIED reference converter (`stage2_accompaniment/convert2midi.py`,
`stage1_compose/convert2midi.py`, `convert_key.degree2pitch`; imported from /root/reference, build container only) run
over a stand-in `miditoolkit` made of plain containers (the real package is not installed), so that the notes / tempo
changes / markers it WOULD hand to miditoolkit are recorded.  The MIDI bytes themselves are not covered.

    python tests/golden/make_midi_golden.py"""
import contextlib, importlib, io, json, os, random, sys, types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("EMO_REFERENCE", "/root/reference")


class Note:
    def __init__(self, velocity=None, pitch=None, start=None, end=None):
        self.velocity, self.pitch, self.start, self.end = velocity, pitch, start, end


class Instrument:
    def __init__(self, program=0, is_drum=False, name=''):
        self.program, self.is_drum, self.name, self.notes = program, is_drum, name, []


class TempoChange:
    def __init__(self, tempo, time):
        self.tempo, self.time = tempo, time


class Marker:
    def __init__(self, text, time):
        self.text, self.time = text, time


class MidiFile:
    def __init__(self):
        self.instruments, self.tempo_changes, self.markers, self.max_tick = [], [], [], 0

    def dump(self, path):
        pass


mt = types.ModuleType("miditoolkit")
mt.midi = types.ModuleType("miditoolkit.midi")
mt.midi.parser = types.ModuleType("miditoolkit.midi.parser")
mt.midi.containers = types.ModuleType("miditoolkit.midi.containers")
mt.midi.parser.MidiFile = MidiFile
for m in (mt, mt.midi.containers):
    m.Note, m.Instrument, m.TempoChange, m.Marker = Note, Instrument, TempoChange, Marker
sys.modules.update({"miditoolkit": mt, "miditoolkit.midi": mt.midi, "miditoolkit.midi.parser": mt.midi.parser,
                    "miditoolkit.midi.containers": mt.midi.containers})


def load(stage_dir, name):
    for k in ("convert2midi", "convert_key"):
        sys.modules.pop(k, None)
    sys.path.insert(0, os.path.join(REF, stage_dir))
    try:
        return importlib.import_module(name)
    finally:
        sys.path.pop(0)


def dump(obj):
    return {"instruments": [[[int(n.velocity), int(n.pitch), int(n.start), int(n.end)] for n in ins.notes] for ins in obj.instruments],
            "tempos": [[int(t.tempo), int(t.time)] for t in obj.tempo_changes],
            "markers": [[str(m.text), int(m.time)] for m in obj.markers], "max_tick": int(obj.max_tick)}


full = ["Tempo_110", "Bar_None", "Beat_0", "Chord_0_M", "Note_Pitch_60", "Note_Duration_480", "Note_Velocity_64",
        "Note_Pitch_64", "Note_Duration_240", "Note_Velocity_70", "Beat_4", "Tempo_96", "Chord_Conti_Conti",
        "Note_Pitch_67", "Note_Duration_120", "Note_Velocity_50", "Beat_8", "Chord_7_7", "Note_Pitch_55",
        "Note_Duration_960", "Note_Velocity_90", "Note_Pitch_50", "Note_Duration_120",      # no velocity: dropped
        "Bar_None", "Beat_0", "Chord_None_None", "Note_Pitch_72", "Note_Duration_1920", "Note_Velocity_100",
        "Beat_12", "Tempo_Conti", "Chord_9_m7", "Note_Pitch_48", "Note_Duration_360", "Note_Velocity_33",
        "Bar_None", "Beat_15", "Chord_9_m7", "Note_Pitch_21", "Note_Duration_120", "Note_Velocity_4", "EOS_None"]
lead = ["Bar_None", "Beat_0", "Chord_0_M", "Note_Pitch_60", "Note_Duration_480", "Beat_4", "Note_Pitch_62",
        "Note_Duration_240", "Beat_8", "Chord_5_M7", "Note_Pitch_65", "Note_Duration_960", "Bar_None", "Beat_0",
        "Chord_5_M7", "Note_Pitch_67", "Note_Duration_480", "Beat_8", "Chord_None_None", "Note_Pitch_69",
        "Note_Duration_480", "Beat_12", "Chord_7_sus4", "Note_Pitch_71", "Note_Duration_480", "Bar_None", "Beat_0",
        "Chord_2_/o7", "Note_Pitch_72", "Note_Duration_1920", "EOS_None"]
out = {"score": [], "degree": []}
with contextlib.redirect_stdout(io.StringIO()):
    c2 = load("stage2_accompaniment", "convert2midi")
    for key in ("Key_C", "Key_F#", "Key_a"):
        out["score"].append({"stage": 2, "key": key, "mode": "full", "play_chords": False, "events": full,
                             "result": dump(c2.event_to_midi(key, full, mode="full"))})
    out["score"].append({"stage": 2, "key": "Key_D", "mode": "skyline", "play_chords": True, "events": lead,
                         "result": dump(c2.event_to_midi("Key_D", lead, mode="skyline", play_chords=True))})
    as_dicts = [{"name": c2.ConversionEvent(e).name, "value": c2.ConversionEvent(e).value} for e in full]
    out["score"].append({"stage": 2, "key": "Key_G", "mode": "full", "play_chords": False, "events": as_dicts, "dict_events": True,
                         "result": dump(c2.event_to_midi("Key_G", as_dicts, mode="full", is_full_event=True))})
    c1 = load("stage1_compose", "convert2midi")
    tempo = [c1.TempoEvent(88, 0, 0)]
    out["score"].append({"stage": 1, "key": "Key_A#", "mode": "lead_sheet", "play_chords": True, "events": lead,
                         "enforce_tempos": [[88, 0]],
                         "result": dump(c1.event_to_midi("Key_A#", lead, mode="lead_sheet", play_chords=True,
                                                         enforce_tempo=True, enforce_tempo_evs=tempo))})
    out["score"].append({"stage": 1, "key": "Key_C", "mode": "full_song", "play_chords": False, "events": full,
                         "result": dump(c1.event_to_midi("Key_C", full, mode="full_song"))})
    for seed in (0, 1, 2, 3, 7):
        random.seed(seed)
        ck = load("stage2_accompaniment", "convert_key")
        keys = list(ck.MAJOR_KEY) + list(ck.MINOR_KEY)
        romans = list(ck.roman2majorDegree)
        out["degree"].append({"seed": seed, "minor_table": {k: int(v) for k, v in ck.roman2minorDegree.items()},
                              "pitch": {"%s|%d|%s" % (k, o, r): int(ck.degree2pitch(k, o, r)) for k in keys for o in (1, 4, 9) for r in romans}})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "midi_score.json"), "w"))
print("scores:", [(s["stage"], s["mode"], len(s["result"]["instruments"]), [len(i) for i in s["result"]["instruments"]],
                   len(s["result"]["markers"])) for s in out["score"]])
print("minor tables:", [(d["seed"], d["minor_table"]["II#"], d["minor_table"]["V#"]) for d in out["degree"]])
