"""On-disk formats of the reference pipeline, so that the B200 path runs on the real EMOPIA+ /
Pop1K7 / HookTheory data and published checkpoints without importing the reference tree:

* `dictionary.pkl` = `(event2idx, idx2event)` written by `representations/events2words.py:88-118`: every observed
  `'{name}_{value}'` event united with a closed-form vocabulary (`build_full_vocab`, :31-85), sorted as strings;
* event pickles: stage 1 `(bar_pos, events)`, stage 2 `(lead_pos, full_pos, events)`
  (`representations/midi2events_emopia.py:761-786`), events being `{'name', 'value'}` dicts;
* generated lead sheets `<name>_<emotion>[_roman].txt`, one event string per line, optional `Key_*` first line
  (`stage2_accompaniment/inference.py:149-166,425-428`).

Host-side, pure python: nothing here is on the measured path.  `tests/test_formats.py` pins the vocabulary and the
dictionary order against the reference functions (`tests/golden/make_formats_golden.py`)."""
import os
import pickle

ROMAN_DEGREES = ('I', 'I#', 'II', 'II#', 'III', 'IV', 'IV#', 'V', 'V#', 'VI', 'VI#', 'VII')   # convert_key.py:33-46
CHORD_QUALITIES = ('M', 'm', 'o', '+', '7', 'M7', 'm7', 'o7', '/o7', 'sus2', 'sus4')             # events2words.py:50
TICKS_PER_16TH, TICKS_PER_BAR = 120, 1920                                                       # events2words.py:8-10


def _int_linspace(lo, hi, n):
    """numpy.linspace(lo, hi, n, dtype=int): truncation of the evenly spaced float grid (last point exact)"""
    step = (hi - lo) / (n - 1)
    return [int(lo + i * step) if i < n - 1 else int(hi) for i in range(n)]


def full_vocab(add_velocity=True, add_emotion=True, add_tempo=True, num_emotion=4, relative=False):
    """the closed-form part of a dictionary (events2words.py:31-85), in the reference's emission order.
    relative=False: REMI (absolute pitches, chord roots as pitch classes 0..11);
    relative=True: functional representation (roman-numeral roots, octave + scale-degree notes)."""
    v = []
    if add_emotion:
        tags = ('Positive', 'Negative', None) if num_emotion == 2 else ('Q1', 'Q2', 'Q3', 'Q4', None)
        v += ['Emotion_%s' % (t,) for t in tags]
    roots = ROMAN_DEGREES if relative else tuple(range(12))
    v += ['Chord_%s_%s' % (r, q) for r in roots for q in CHORD_QUALITIES]
    v.append('Chord_None_None')
    if relative:
        v += ['Note_Octave_%d' % o for o in range(21 // 12, 109 // 12 + 1)]
        v += ['Note_Degree_%s' % d for d in ROMAN_DEGREES]
    else:
        v += ['Note_Pitch_%d' % p for p in range(21, 109)]
    if add_velocity:
        v += ['Note_Velocity_%d' % x for x in _int_linspace(4, 127, 42)]
    v += ['Note_Duration_%d' % d for d in range(TICKS_PER_16TH, TICKS_PER_BAR + TICKS_PER_16TH, TICKS_PER_16TH)]
    if add_tempo:
        v += ['Tempo_%d' % t for t in _int_linspace(32, 224, 65)]
    return v


# the flag sets the reference builds its six dictionaries with (events2words.py:139-171)
VOCAB_FLAGS = {
    'stage1_lead_sheet': dict(add_velocity=False, add_emotion=True, add_tempo=False, num_emotion=2),
    'stage2_full_song': dict(add_velocity=True, add_emotion=True, add_tempo=True, num_emotion=4),
    'one_stage_full_song': dict(add_velocity=True, add_emotion=True, add_tempo=True, num_emotion=4),
}


def event_name(e):
    """'Name_Value' of an event stored as a dict (pickles) or already as a string (lead-sheet text)"""
    return '%s_%s' % (e['name'], e['value']) if isinstance(e, dict) else e


def build_dictionary(event_seqs, relative=False, **flags):
    """(event2idx, idx2event) over the observed events of `event_seqs` united with the closed-form vocabulary,
    ids in sorted string order (events2dictionary, events2words.py:88-113)."""
    seen = set(full_vocab(relative=relative, **flags))
    for seq in event_seqs:
        seen.update(event_name(e) for e in seq)
    names = sorted(seen)
    return {n: i for i, n in enumerate(names)}, {i: n for i, n in enumerate(names)}


def save_dictionary(path, event2idx, idx2event):
    with open(path, 'wb') as f:
        pickle.dump((event2idx, idx2event), f)


def load_dictionary(path):
    """-> (event2idx, idx2event, vocab_size) with vocab_size counting the PAD id the models append
    (stage2 dataloader.py:66-74: pad_token = len(event2idx), vocab_size = pad_token + 1)"""
    with open(path, 'rb') as f:
        event2idx, idx2event = pickle.load(f)[:2]
    return event2idx, idx2event, len(event2idx) + 1


def save_piece(path, *fields):
    """stage 1: save_piece(p, bar_pos, events); stage 2: save_piece(p, lead_pos, full_pos, events)"""
    with open(path, 'wb') as f:
        pickle.dump(tuple(fields), f)


def load_piece(path):
    """-> dict(bar_pos | lead_pos + full_pos, events) according to the tuple length stored"""
    with open(path, 'rb') as f:
        t = pickle.load(f)
    if len(t) == 2:
        return {'bar_pos': t[0], 'events': t[1]}
    return {'lead_pos': t[0], 'full_pos': t[1], 'events': t[2]}


def piece_files(data_dir, piece_names=None):
    """the pickles of a split: `pieces` is the list stored in the train / valid split pickle (file names)"""
    names = sorted(os.listdir(data_dir)) if piece_names is None else list(piece_names)
    return [os.path.join(data_dir, n) for n in names if n.endswith('.pkl')]


def read_lead_sheet(path, event2idx):
    """generated lead sheet -> (key event, list of bars of token ids).  The key defaults to 'Key_C' when the first
    line is not a key event; everything before the first 'Bar_None' (the key / emotion header) is dropped
    (inference.py:149-166)."""
    with open(path) as f:
        events = f.read().splitlines()
    key = events[0] if events and 'Key' in events[0] else 'Key_C'
    starts = [i for i, e in enumerate(events) if e == 'Bar_None'] + [len(events)]
    bars = [[event2idx[e] for e in events[a:b]] for a, b in zip(starts[:-1], starts[1:])]
    return key, bars


def write_events(path, events):
    """one event string per line (what both inference scripts write next to the .mid)"""
    with open(path, 'w') as f:
        for e in events:
            f.write(event_name(e) + '\n')


def lead_sheet_files(directory, representation):
    """the stage-1 outputs stage 2 consumes: '*roman.txt' for the functional representation, any '.txt' otherwise
    (inference.py:425-428); accompaniments already written ('_full') are not lead sheets"""
    tag = 'roman.txt' if representation == 'functional' else '.txt'
    return sorted(os.path.join(directory, f) for f in os.listdir(directory) if tag in f and '_full' not in f)


def emotions_for(file_name):
    """lead-sheet file name -> the quadrants to render (inference.py:433-450)"""
    for tag, quads in (('Positive', ['Q1', 'Q4']), ('Negative', ['Q2', 'Q3']), ('Q1', ['Q1']), ('Q2', ['Q2']),
                       ('Q3', ['Q3']), ('Q4', ['Q4']), ('None', ['None'])):
        if tag in file_name:
            return quads
    raise ValueError('wrong emotion label')
