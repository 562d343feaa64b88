"""Event sequence -> MIDI file (second half): the post-processing both inference scripts end with,
without `miditoolkit` (absent in this image; the reference's `convert2midi.py` cannot even be imported without it).

* `relative_to_absolute`   functional events -> absolute pitches / pitch-class chord roots
                           (stage1_compose/inference.py:44-72, stage2_accompaniment/inference.py:172-198, convert_key.py:143-155)
* `full_track_bars`        the Full-track events of every generated bar (stage2 inference.py:200-208)
* `events_to_score`        notes / tempo changes / chord + bar markers in ticks, optional block chords
                           (stage*/convert2midi.py:149-300; modes 'full' = 'full_song', 'skyline' = 'lead_sheet')
* `write_midi`, `read_midi`  Standard MIDI File, format 1, 480 ticks per beat: track 0 = tempo + markers, one track per
                           instrument (what `MidiFile.dump` produces; byte-level identity with miditoolkit is NOT claimed --
                           parity unpinned for the file bytes, the score is pinned: tests/golden/make_midi_golden.py runs
                           the UNMODIFIED reference converter over a stand-in container module)

Host-side, pure python; nothing here is on the measured path."""
import random
import struct

TICKS_PER_BEAT, TICKS_PER_BAR, POSITIONS_PER_BAR = 480, 1920, 16
SHARP_NAMES = ('C', 'C#', 'D', 'D#', 'E', 'F', 'F#', 'G', 'G#', 'A', 'A#', 'B')                # convert2midi.py:13
MAJOR_DEGREE = {'I': 0, 'I#': 1, 'II': 2, 'II#': 3, 'III': 4, 'IV': 5, 'IV#': 6, 'V': 7, 'V#': 8, 'VI': 9,
                'VI#': 10, 'VII': 11}                                                               # convert_key.py:33-47
QUALITY_INTERVALS = {                                                                                # convert2midi.py:17-52,72-85
    'M': (0, 4, 7), 'm': (0, 3, 7), '+': (0, 4, 8), 'o': (0, 3, 6), 'sus4': (0, 5, 7), 'sus2': (0, 2, 7),
    '7': (0, 4, 7, 10), 'M7': (0, 4, 7, 11), 'm7': (0, 3, 7, 10), 'o7': (0, 3, 6, 9), '/o7': (0, 3, 6, 10),
}


def minor_degree_table(rng=random):
    """roman numeral -> semitones above the tonic in a minor key.  The reference draws the two ambiguous entries with
    `random.choice` when `convert_key` is imported (convert_key.py:49-76: two draws for the inverse table first, then
    'II#' in (2, 3) and 'V#' in (7, 8)); the same four draws are made here so a seeded process agrees with it."""
    rng.choice(['III', 'IV'])
    rng.choice(['VII', 'I'])
    return {'I': 0, 'I#': 1, 'II': 2, 'II#': rng.choice([2, 3]), 'III': 3, 'IV': 5, 'IV#': 6, 'V': 7,
            'V#': rng.choice([7, 8]), 'VI': 8, 'VI#': 9, 'VII': 10}


MINOR_DEGREE = minor_degree_table()


def degree_to_pitch(keyname, octave, roman, minor_degree=None):
    """convert_key.py:143-155: upper-case key names are major keys, lower-case minor"""
    if keyname in SHARP_NAMES:
        return octave * 12 + SHARP_NAMES.index(keyname) + MAJOR_DEGREE[roman]
    if keyname.upper() in SHARP_NAMES and keyname == keyname.lower():
        return octave * 12 + SHARP_NAMES.index(keyname.upper()) + (minor_degree or MINOR_DEGREE)[roman]
    raise NameError('Wrong key name {}.'.format(keyname))


def relative_to_absolute(key, events, keep_conti_chords=True, minor_degree=None):
    """'Note_Octave_o' + 'Note_Degree_r' -> 'Note_Pitch_p' (clamped to the piano range 21..108), 'Chord_<roman>_<q>' ->
    'Chord_<semitones above the tonic>_<q>'; everything else is passed through.  key = 'Key_<name>'."""
    keyname = key.split('_')[1]
    table = MAJOR_DEGREE if keyname in SHARP_NAMES else (minor_degree or MINOR_DEGREE)
    out, octave = [], None
    for ev in events:
        if 'Note_Octave' in ev:
            octave = int(ev.split('_')[2])
        elif 'Note_Degree' in ev:
            pitch = degree_to_pitch(keyname, octave, ev.split('_')[2], minor_degree)
            out.append('Note_Pitch_%d' % min(108, max(21, pitch)))
        elif 'Chord_' in ev:
            if 'None' in ev or (keep_conti_chords and 'Conti' in ev):
                out.append(ev)
            else:
                _, root, quality = ev.split('_')[:3]
                out.append('Chord_%d_%s' % (table[root], quality))
        else:
            out.append(ev)
    return out


def full_track_bars(events):
    """per generated bar, the events between 'Track_Full' and the next 'Track_LeadSheet' (or the end)"""
    lead = [i for i, e in enumerate(events) if e == 'Track_LeadSheet']
    full = [i for i, e in enumerate(events) if e == 'Track_Full']
    return [events[st + 1:ed] for st, ed in zip(full, lead[1:] + [len(events)])]


def _name_value(ev):
    if isinstance(ev, dict):
        return ev['name'], ev['value']
    parts = ev.split('_')
    if 'Note' in ev:
        return '_'.join(parts[:-1]), parts[-1]
    if 'Chord' in ev:
        return parts[0], '_'.join(parts[1:])
    name, value = parts                      # any other event has exactly one '_' (the reference unpacks two values)
    return name, value


def chord_pitches(chord):
    """'<root name>_<quality>' -> [bass in octave 2] + chord tones above middle C (convert2midi.py:290-303)"""
    root, quality = chord.split('_')
    pc = SHARP_NAMES.index(root)
    return [36 + pc] + [60 + pc + i for i in QUALITY_INTERVALS[quality]]


def events_to_score(key, events, mode='full', play_chords=False, enforce_tempos=None):
    """-> dict(instruments=[[(velocity, pitch, start, end), ...], ...], tempos=[(bpm, tick)], markers=[(text, tick)],
    max_tick).  mode 'full' / 'full_song': a note is Pitch, Duration, Velocity; 'skyline' / 'lead_sheet': Pitch, Duration
    at velocity 80.  enforce_tempos: [(bpm, tick)] replacing the tempo events of the sequence (stage 1 passes the prompt
    tempo); play_chords adds a second piano track with the chord markers rendered as block chords at velocity 63."""
    keyname = key.split('_')[1].upper()
    start = SHARP_NAMES.index(keyname)
    evs = [_name_value(e) for e in events]
    full = mode in ('full', 'full_song')
    lead = mode in ('skyline', 'lead_sheet')
    notes, tempos, chords = [], [], []
    bar, pos = -1, 0
    tick = lambda: bar * TICKS_PER_BAR + pos * (TICKS_PER_BAR // POSITIONS_PER_BAR)
    for i, (name, value) in enumerate(evs):
        if name == 'Bar':
            bar += 1
        elif name == 'Beat':
            pos = int(value)
            assert 0 <= pos < POSITIONS_PER_BAR
        elif name == 'Tempo' and 'Conti' not in str(value):
            tempos.append((int(value), max(bar, 0) * TICKS_PER_BAR + pos * (TICKS_PER_BAR // POSITIONS_PER_BAR)))
        elif 'Note_Pitch' in name:
            nxt = evs[i + 1][0] if i + 1 < len(evs) else ''
            nxt2 = evs[i + 2][0] if i + 2 < len(evs) else ''
            if full and 'Note_Duration' in nxt and 'Note_Velocity' in nxt2:
                notes.append((int(evs[i + 2][1]), int(value), tick(), tick() + int(evs[i + 1][1])))
            elif lead and 'Note_Duration' in nxt:
                notes.append((80, int(value), tick(), tick() + int(evs[i + 1][1])))
        elif 'Chord' in name and 'Conti' not in str(value):
            chords.append((str(value), tick()))
    markers = []
    for val, t in chords:
        if 'None' not in val:                # root = semitones above the tonic -> note name in the piece's key
            root, quality = val.split('_')[:2]
            val = SHARP_NAMES[(start + int(root)) % 12] + '_' + quality
        markers.append(('Chord-' + val, t))
    markers += [('Bar-%d' % (b + 1), TICKS_PER_BAR * b) for b in range(bar)]
    score = {'instruments': [notes], 'tempos': list(enforce_tempos) if enforce_tempos is not None else tempos,
             'markers': markers, 'max_tick': max(n[3] for n in notes)}
    if play_chords:
        uniq, prev = [], None
        for text, t in markers:
            if not text.startswith('Chord') or text == 'Chord-None_None':
                continue
            if text != prev:
                prev = text
                uniq.append((text, t))
        block = []
        for (text, t), nxt in zip(uniq, [u[1] for u in uniq[1:]] + [score['max_tick']]):
            block += [(63, p, t, nxt) for p in chord_pitches(text.split('-')[1])]
        score['instruments'].append(block)
    return score


# ------------------------------------------------------------------------------------------------------------
# Standard MIDI File writer / reader
# ------------------------------------------------------------------------------------------------------------
def _vlq(n):
    if n < 0:
        raise ValueError("negative tick / delta time (an event before the first 'Bar' of the sequence?)")
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    return bytes(reversed(out))


def _track(events):
    """events: (tick, order, bytes) -> MTrk chunk with delta times and an end-of-track meta event"""
    body, last = bytearray(), 0
    for t, _, data in sorted(events, key=lambda e: (e[0], e[1])):
        body += _vlq(t - last) + data
        last = t
    body += b'\x00\xff\x2f\x00'
    return b'MTrk' + struct.pack('>I', len(body)) + bytes(body)


def _meta(kind, payload):
    return bytes([0xFF, kind]) + _vlq(len(payload)) + payload


def write_midi(path, score, instrument_name='Piano'):
    tracks = []
    ev = [(t, 0, _meta(0x51, struct.pack('>I', int(round(60_000_000 / bpm)))[1:])) for bpm, t in score['tempos']]
    ev += [(t, 1, _meta(0x06, text.encode('utf-8'))) for text, t in score['markers']]
    tracks.append(_track(ev))
    for notes in score['instruments']:
        ev = [(0, 0, _meta(0x03, instrument_name.encode())), (0, 1, bytes([0xC0, 0]))]
        for vel, pitch, st, ed in notes:
            ev.append((int(st), 3, bytes([0x90, int(pitch) & 0x7F, int(vel) & 0x7F])))
            ev.append((int(ed), 2, bytes([0x80, int(pitch) & 0x7F, 0])))          # note-offs first at equal ticks
        tracks.append(_track(ev))
    with open(path, 'wb') as f:
        f.write(b'MThd' + struct.pack('>IHHH', 6, 1, len(tracks), TICKS_PER_BEAT))
        for tr in tracks:
            f.write(tr)


def read_midi(path):
    """minimal SMF reader (what write_midi emits + running status): -> the same dict layout as events_to_score"""
    data = open(path, 'rb').read()
    assert data[:4] == b'MThd'
    _, fmt, ntrk, div = struct.unpack('>IHHH', data[4:14])
    pos, tempos, markers, instruments = 14, [], [], []
    for _ in range(ntrk):
        assert data[pos:pos + 4] == b'MTrk'
        ln = struct.unpack('>I', data[pos + 4:pos + 8])[0]
        p, end, t, status = pos + 8, pos + 8 + ln, 0, 0
        pos = end
        on, notes, has_notes = {}, [], False
        while p < end:
            d = 0
            while True:
                b = data[p]; p += 1
                d = (d << 7) | (b & 0x7F)
                if not b & 0x80:
                    break
            t += d
            if data[p] & 0x80:
                status = data[p]; p += 1
            if status == 0xFF:
                kind = data[p]; p += 1
                n = 0
                while True:
                    b = data[p]; p += 1
                    n = (n << 7) | (b & 0x7F)
                    if not b & 0x80:
                        break
                payload = data[p:p + n]; p += n
                if kind == 0x51:
                    tempos.append((int(round(60_000_000 / int.from_bytes(payload, 'big'))), t))
                elif kind == 0x06:
                    markers.append((payload.decode('utf-8'), t))
            elif status & 0xF0 in (0x80, 0x90):
                pitch, vel = data[p], data[p + 1]; p += 2
                has_notes = True
                if status & 0xF0 == 0x90 and vel > 0:
                    on.setdefault(pitch, []).append((t, vel))
                elif on.get(pitch):
                    st, v = on[pitch].pop(0)
                    notes.append((v, pitch, st, t))
            elif status & 0xF0 in (0xC0, 0xD0):
                p += 1
            else:
                p += 2
        if has_notes:
            instruments.append(notes)
    return {'ticks_per_beat': div, 'format': fmt, 'instruments': instruments, 'tempos': tempos, 'markers': markers,
            'max_tick': max((n[3] for ins in instruments for n in ins), default=0)}
