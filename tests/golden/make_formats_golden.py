"""Mint tests/golden/formats.json from the UNMODIFIED reference vocabulary builder
(`representations/events2words.py`, imported from /root/reference; build container only): the closed-form vocabulary
for every flag set the reference uses, and a dictionary built by `events2dictionary` over synthetic event pickles in
the reference's on-disk layout (stage-2 tuples of three, dict-typed events).

    python tests/golden/make_formats_golden.py"""
import contextlib, io, json, os, pickle, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("EMO_REFERENCE", "/root/reference")
sys.modules.setdefault("pickle5", pickle)
sys.path.insert(0, os.path.join(REF, "representations"))
import events2words as ref                                      # noqa: E402  (the reference module, unmodified)

out = {"vocab": [], "dictionary": {}}
with contextlib.redirect_stdout(io.StringIO()):
    for relative in (False, True):
        for vel, emo, tempo, nemo in ((False, True, False, 2), (True, True, True, 4), (True, False, True, 4), (False, True, True, 2)):
            v = ref.build_full_vocab(add_velocity=vel, add_emotion=emo, add_tempo=tempo, num_emotion=nemo, relative=relative)
            out["vocab"].append({"relative": relative, "add_velocity": vel, "add_emotion": emo, "add_tempo": tempo,
                                 "num_emotion": nemo, "events": v})

# a dictionary over observed events: structural events the closed form does not contain, mixed value types
observed = [
    [{"name": "Emotion", "value": "Q1"}, {"name": "Key", "value": "C"}, {"name": "Tempo", "value": 110},
     {"name": "Bar", "value": None}, {"name": "Beat", "value": 0}, {"name": "Track", "value": "LeadSheet"},
     {"name": "Note_Octave", "value": 4}, {"name": "Note_Degree", "value": "V"}, {"name": "Note_Duration", "value": 480},
     {"name": "Beat", "value": 12}, {"name": "Track", "value": "Full"}, {"name": "Note_Velocity", "value": 64},
     {"name": "EOS", "value": None}],
    [{"name": "Key", "value": "a"}, {"name": "Bar", "value": None}, {"name": "Beat", "value": 3},
     {"name": "Chord", "value": "IV_M7"}, {"name": "Tempo", "value": 33}, {"name": "EOS", "value": None}],
]
with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "tests", "golden")) as tmp:
    os.makedirs(os.path.join(tmp, "events"))
    for i, ev in enumerate(observed):
        pickle.dump(([[0, 3]], [[0, 5]], ev), open(os.path.join(tmp, "events", "p%d.pkl" % i), "wb"))
    with contextlib.redirect_stdout(io.StringIO()):
        ref.events2dictionary(tmp, add_velocity=True, add_emotion=True, add_tempo=True, num_emotion=4, relative=True, event_pos=2)
    e2i, i2e = pickle.load(open(os.path.join(tmp, "dictionary.pkl"), "rb"))
out["dictionary"] = {"observed": observed, "event2idx": e2i, "idx2event": {str(k): v for k, v in i2e.items()},
                     "flags": {"add_velocity": True, "add_emotion": True, "add_tempo": True, "num_emotion": 4, "relative": True}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "formats.json"), "w"), indent=0)
print("vocab sets:", [len(v["events"]) for v in out["vocab"]], "| dictionary:", len(e2i))
