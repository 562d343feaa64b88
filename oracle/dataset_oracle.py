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
