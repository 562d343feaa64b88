"""CPU oracle: stage-2 dataset sample assembly.  TEST INFRASTRUCTURE.

Restates, in numpy, what `REMISkylineToMidiTransformerDataset.__getitem__` does for one (piece, start bar):
  reference stage2_accompaniment/dataloader.py
    :29-39    convert_event            events ('Name_Value' strings or {'name','value'} dicts) -> ids
    :83-106   build_dataset            admissible start bars of a piece
    :117-125  pad_sequence
    :127-145  make_target_and_mask     targets = next token inside the Full-track span of every bar from st_bar on
                                       (EOS closes the last bar), PAD elsewhere; track_mask = 1 on those spans
    :147-173  make_target_and_mask_predict   (+ emotion / key positions)
    :178-231  __getitem__              header + events from the start bar, pad / truncate to model_dec_seqlen,
                                       chord_idx / melody_idx = the TARGET event is a 'Chord_*' / 'Note_*' event
Pinned against the unmodified reference class on synthetic pieces by tests/golden/make_dataset_golden.py (run in the
build container, fixtures committed as tests/golden/dataset_small.npz)."""
import numpy as np


def event_name(e):
    return '{}_{}'.format(e['name'], e['value']) if isinstance(e, dict) else e


def admissible_stbars(n_events, melody_pos, seqlen):
    """dataloader.py:95-106"""
    if n_events <= seqlen:
        return [0]
    out = []
    for bar in range(len(melody_pos)):
        if n_events - melody_pos[bar][0] >= 0.5 * seqlen:
            out.append(bar)
        else:
            break
    return out


def assemble(piece_tokens, melody_pos, chord_pos, st_bar, seqlen, pad_token, eos_token, is_chord, is_note,
             predict_key=False):
    """piece_tokens: ids of ALL events of the piece.  Returns dict(dec_input, dec_target, track_mask, chord_idx,
    melody_idx [seqlen] int64, length)."""
    hdr = melody_pos[0][0]
    toks = list(piece_tokens[:hdr]) + list(piece_tokens[melody_pos[st_bar][0]:])          # :186-190
    length = len(toks)
    inp = toks + [pad_token] * max(0, seqlen - len(toks))                                   # :194-197 (pad_to_same)
    inp = np.array(inp, dtype=np.int64)
    tgt = np.full_like(inp, pad_token)
    mask = np.zeros_like(inp)
    if predict_key:                                                                          # :153-158
        mask[0], mask[1] = 2, 3
        tgt[0] = inp[1]
    off = -melody_pos[st_bar][0] + melody_pos[0][0]
    nb = len(melody_pos)
    for b in range(st_bar, nb):                                                              # :131-143
        c0, c1 = chord_pos[b][0] + off, chord_pos[b][1] + off
        mask[c0:c1] = 1
        if b != nb - 1:
            tgt[c0:c1] = inp[c0 + 1:c1 + 1]
        else:
            tgt[c0:c1 - 1] = inp[c0 + 1:c1]
            tgt[c1 - 1] = eos_token
    chord_idx = np.asarray(is_chord)[tgt].astype(np.int64)                                   # :206-213
    melody_idx = np.asarray(is_note)[tgt].astype(np.int64)
    sl = slice(0, seqlen)                                                                    # :215-219
    return {"dec_input": inp[sl], "dec_target": tgt[sl], "track_mask": mask[sl], "chord_idx": chord_idx[sl],
            "melody_idx": melody_idx[sl], "length": min(length, seqlen)}


def vocab_flags(idx2event, pad_token):
    """per-id 'the event is a Chord_* / Note_* event' (PAD = 'Pad_None': neither), dataloader.py:206-213"""
    V = pad_token + 1
    is_chord, is_note = np.zeros(V, dtype=np.int64), np.zeros(V, dtype=np.int64)
    for i in range(pad_token):
        t = idx2event[i].split('_')[0]
        is_chord[i] = t == 'Chord'
        is_note[i] = t == 'Note'
    return is_chord, is_note


def synthetic_piece(event2idx, n_bars, rng, lead_len=(3, 9), full_len=(5, 40), as_dicts=False):
    """A piece in the on-disk format of representations/*/midi2events (melody_pos, chord_pos, events): a 3-event
    header (emotion, key, tempo), then per bar a lead-sheet span followed by a full-track span."""
    names = [e for e in event2idx if e.split('_')[0] in ('Note', 'Chord', 'Beat')]
    ev = ['Emotion_Q%d' % rng.randint(1, 5), 'Key_C', 'Tempo_110']
    mel, ch = [], []
    for _ in range(n_bars):
        s = len(ev)
        ev += ['Track_LeadSheet', 'Bar_None'] + [names[i] for i in rng.randint(0, len(names), rng.randint(*lead_len))]
        m = len(ev)
        ev += ['Track_Full'] + [names[i] for i in rng.randint(0, len(names), rng.randint(*full_len))]
        mel.append((s, m))
        ch.append((m, len(ev)))
    if as_dicts:
        ev = [{'name': e.split('_')[0], 'value': '_'.join(e.split('_')[1:])} for e in ev]
    return mel, ch, ev


# ---------------------------------------------------------------------------------------------------------
# stage-1 dataset: SkylineFullSongTransformerDataset as stage1_compose/train.py:230-262 builds it
# (max_n_seg from the YAML = 1, do_augment=False): one sample = the FIRST segment of a piece.
#   reference stage1_compose/dataloader.py
#     :354-384  build_dataset        bar positions: appended marker / trailing empty bar removed, sentinel appended
#     :386-406  register_segments    first segment = bars [0, b) fitting model_dec_seqlen
#     :408-445  get_sample_from_file events up to the last registered bar + closing EOS_None / Bar_None
#     :469-520  get_decoder_input_data   input / shifted target / Chord- and Note-target masks of the segment
#     :194-255  collate_fn           PAD-fill to model_dec_seqlen (the two masks are PAD-filled as well)
# The encoder-side features (enc_inp, chroma, groove, masks: :533-608) are never read by train.py and are not restated.
# Pinned against the unmodified class by tests/golden/make_stage1_dataset_golden.py (tests/golden/stage1_dataset.npz).
# ---------------------------------------------------------------------------------------------------------
def stage1_bar_positions(bar_pos, n_events, max_bars):
    """:354-384 -> (registered bar positions incl. the sentinel, sample closes with EOS (else Bar))"""
    bp, n = list(bar_pos), n_events
    if bp[-1] == n:                       # appended bar marker
        bp = bp[:-1]
    if n - bp[-1] == 2:                   # trailing empty bar: [Bar_None, EOS_None]
        n = bp[-1]
        bp = bp[:-1]
    if len(bp) <= max_bars:
        bp.append(n - 1)                  # the position of <EOS>
    else:
        bp = bp[:max_bars + 1]
    return bp, len(bp) <= max_bars        # :431 tests the list WITH the sentinel


def stage1_first_segment(bp, seqlen):
    """:386-406 with the break after the first cut: (start bar, end bar) of segment 0"""
    for b in range(len(bp) - 1):
        if bp[b + 1] - bp[0] > seqlen - 1 and b > 0:
            return 0, b
    return 0, len(bp) - 1


def stage1_sample_tokens(piece_tokens, bp, closes_with_eos, eos_token, bar_token):
    """:408-437: ids of events[: bp[-1]] + the closing token"""
    return list(piece_tokens[:bp[-1]]) + [eos_token if closes_with_eos else bar_token]


def stage1_assemble(piece_tokens, bar_pos, seqlen, max_bars, pad_token, eos_token, bar_token, is_chord, is_note):
    """one collated row: dict(dec_inp, dec_tgt, inp_chord, inp_melody [seqlen] int64, dec_seg_len)"""
    bp, eos = stage1_bar_positions(bar_pos, len(piece_tokens), max_bars)
    toks = stage1_sample_tokens(piece_tokens, bp, eos, eos_token, bar_token)
    s, e = stage1_first_segment(bp, seqlen)
    E = bp[e] - bp[s] + 1                                           # :483-484 (segment offsets are relative to bp[0],
    inp, tgt = toks[0:E], toks[1:E + 1]                             #           the slices start at token 0)
    if len(inp) != len(tgt):
        raise AssertionError("segment runs past the sample (the reference asserts, :512)")
    seg_len = len(inp)                                              # recorded before truncation (:492)
    tgt_a = np.asarray(tgt[:seqlen], dtype=np.int64)
    out = {k: np.full(seqlen, pad_token, dtype=np.int64) for k in ("dec_inp", "dec_tgt", "inp_chord", "inp_melody")}
    n = len(tgt_a)
    out["dec_inp"][:n] = inp[:seqlen]
    out["dec_tgt"][:n] = tgt_a
    out["inp_chord"][:n] = np.asarray(is_chord)[tgt_a]
    out["inp_melody"][:n] = np.asarray(is_note)[tgt_a]
    out["dec_seg_len"] = seg_len
    return out


def synthetic_stage1_piece(rng, n_bars, bar_len=(3, 14), tail="eos", relative=True):
    """(bar_pos, events) in the stage-1 pickle layout: [Emotion, Key] header, bars of Beat / Chord / Note events.
    tail: 'eos' (.. EOS), 'empty_bar' (.. Bar, EOS), 'marker' (bar_pos has len(events) appended)."""
    deg = ('I', 'II', 'III', 'IV', 'V', 'VI', 'VII')
    ev = [{'name': 'Emotion', 'value': ('Positive', 'Negative')[rng.randint(2)]}, {'name': 'Key', 'value': 'C'}]
    bar_pos = []
    for _ in range(n_bars):
        bar_pos.append(len(ev))
        ev.append({'name': 'Bar', 'value': None})
        n, beat = rng.randint(*bar_len), 0
        while n > 0:
            ev.append({'name': 'Beat', 'value': beat})
            ev.append({'name': 'Chord', 'value': '%s_%s' % (deg[rng.randint(7)], ('M', 'm', '7')[rng.randint(3)])})
            if relative:
                ev.append({'name': 'Note_Octave', 'value': int(rng.randint(3, 7))})
                ev.append({'name': 'Note_Degree', 'value': deg[rng.randint(7)]})
            else:
                ev.append({'name': 'Note_Pitch', 'value': int(rng.randint(48, 84))})
            ev.append({'name': 'Note_Duration', 'value': 120 * int(rng.randint(1, 9))})
            n -= 5
            beat = min(15, beat + int(rng.randint(1, 5)))
    if tail == "empty_bar":
        bar_pos.append(len(ev))
        ev.append({'name': 'Bar', 'value': None})
    ev.append({'name': 'EOS', 'value': None})
    if tail == "marker":
        bar_pos.append(len(ev))
    return bar_pos, ev
