"""Synthetic stand-ins for the data the reference reads from disk (no dataset / dictionary.pkl / checkpoint
is reachable offline): a vocabulary with the same structural events as `representations/events2words.py`
builds (Bar / Beat_0..15 / Track_* / EOS / Emotion_* / Key_* / Tempo_* / Note_* / Chord_* ... + PAD last),
random token batches shaped like `dataloader.py` batches, and random lead sheets."""
import numpy as np
import torch

MAJOR = ['C', 'C#', 'D', 'D#', 'E', 'F', 'F#', 'G', 'G#', 'A', 'A#', 'B']


def synthetic_vocab(V, stage=2):
    """(event2idx, idx2event) with exactly V entries; PAD_None is V-1 (the models' ignore_index)."""
    ev = ['Bar_None', 'EOS_None'] + ['Beat_%d' % b for b in range(16)]
    if stage == 2:
        ev += ['Track_LeadSheet', 'Track_Full'] + ['Emotion_Q%d' % q for q in range(1, 5)] + ['Emotion_None']
    else:
        ev += ['Emotion_Positive', 'Emotion_Negative']
    ev += ['Key_%s' % k for k in MAJOR] + ['Key_%s' % k.lower() for k in MAJOR]
    ev += ['Tempo_Conti'] + ['Tempo_%d' % t for t in range(50, 200, 20)]          # includes Tempo_110
    ev += ['Chord_Conti_Conti', 'Chord_None_None']
    if V - 1 < len(ev):
        raise ValueError('synthetic vocabulary needs V >= %d' % (len(ev) + 1))
    i = 0
    kinds = ['Note_Octave_%d', 'Note_Degree_%d', 'Note_Duration_%d', 'Note_Velocity_%d', 'Note_Pitch_%d', 'Chord_I_%d']
    seen = set(ev)
    while len(ev) < V - 1:
        e = kinds[i % len(kinds)] % (i // len(kinds))
        if e not in seen:
            ev.append(e)
            seen.add(e)
        i += 1
    ev = ev[:V - 1] + ['PAD_None']
    assert len(ev) == V and len(set(ev)) == V
    event2idx = {e: i for i, e in enumerate(ev)}
    idx2event = {i: e for i, e in enumerate(ev)}
    return event2idx, idx2event


def synthetic_batch(V, B, T, seed):
    """tokens uniform over [0, V-2], seg ~ Bernoulli(0.5), targets = inputs shifted by one with PAD (= V-1)
    wherever the position is not on the Full track (mimics stage2 dataloader.py:127-144)."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, V - 1, (B, T), generator=g)
    seg = torch.randint(0, 2, (B, T), generator=g)
    tgt = torch.where(seg == 1, torch.roll(tok, -1, 1), torch.full_like(tok, V - 1))
    return tok, seg, tgt


def synthetic_lead_sheet(event2idx, n_bars, seed, events_per_bar=12):
    """list of bars, each [Bar_None, Beat_k, notes...] as token ids (what read_generated_events returns)."""
    rng = np.random.RandomState(seed)
    notes = [i for e, i in event2idx.items() if e.startswith('Note_') or e.startswith('Chord_')]
    bars = []
    for _ in range(n_bars):
        bar = [event2idx['Bar_None']]
        beats = sorted(rng.choice(16, size=3, replace=False).tolist())
        per = max(1, (events_per_bar - 1 - len(beats)) // len(beats))
        for b in beats:
            bar.append(event2idx['Beat_%d' % b])
            bar.extend(int(x) for x in rng.choice(notes, size=per))
        bars.append(bar)
    return bars
