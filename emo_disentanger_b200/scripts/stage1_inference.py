"""Stage-1 generation -- the reference's `python3 stage1_compose/inference.py -c CONFIG -r REPR -m MODE [-i PARAMS]
[-o OUT_DIR] [-n N_GROUPS]` surface (inference.py:86-298): builds the model with mem_len = tgt_len, generates
one lead sheet per valence (Positive / Negative) per group with t = 1.2, p = 0.97 (lead_sheet) and writes the
`samp_XX_<Emotion>[_roman].txt` event files stage 2 consumes, the absolute-pitch `.txt` and the `.mid`
(tempo 110, block chords under the melody in lead-sheet mode, inference.py:255-283) through `data/midi_out.py`."""
import argparse
import os
import shutil
import numpy as np
import torch
import yaml

from ..generate import generate_plain_xl
from ..synth import synthetic_vocab
from ..data.formats import load_dictionary, write_events
from ..data.midi_out import relative_to_absolute, events_to_score, write_midi
from .stage1_train import build_model


def write_lead_sheet_outputs(out_dir, out_name, events, rep, mode):
    """what the reference leaves per generated piece (inference.py:252-283): `<name>_roman.txt` (functional events),
    `<name>.txt` (absolute events) and `<name>.mid` (tempo 110; lead-sheet mode adds the block-chord track).
    events[0] is the emotion primer and is dropped from all three.  Returns the files written."""
    written = []
    key = next((e for e in events if 'Key' in e), 'Key_C')
    if rep == 'functional':
        write_events(os.path.join(out_dir, out_name + '_roman.txt'), events[1:])
        written.append(out_name + '_roman.txt')
        try:
            events = relative_to_absolute(key, events, keep_conti_chords=False)
        except (KeyError, TypeError, ValueError) as e:     # random weights: a degree before any octave event, ...
            print('[info] {}: functional -> absolute conversion failed ({}); roman events only'.format(out_name, type(e).__name__))
            return written
    write_events(os.path.join(out_dir, out_name + '.txt'), events[1:])
    written.append(out_name + '.txt')
    try:
        score = events_to_score(key, events[1:], mode=mode, play_chords=(mode == 'lead_sheet'),
                                enforce_tempos=[(110, 0)] if mode == 'lead_sheet' else None)
        write_midi(os.path.join(out_dir, out_name + '.mid'), score)
        written.append(out_name + '.mid')
    except (ValueError, AssertionError, KeyError) as e:    # e.g. no complete note in the sequence
        print('[info] {}: no MIDI written ({})'.format(out_name, type(e).__name__))
    return written


def main(argv=None):
    ap = argparse.ArgumentParser(description='')
    req = ap.add_argument_group('required arguments')
    req.add_argument('-c', '--configuration', required=True)
    req.add_argument('-r', '--representation', choices=['remi', 'functional'], required=True)
    req.add_argument('-m', '--mode', choices=['lead_sheet', 'full_song'], required=True, help='generation mode')
    ap.add_argument('-i', '--inference_params', default=None)
    ap.add_argument('-o', '--output_dir', default='generation/emopia_functional_two')
    ap.add_argument('-p', '--play_midi', default=False, action='store_true')
    ap.add_argument('-n', '--n_groups', default=20)
    ap.add_argument('--synthetic', type=int, default=0)
    ap.add_argument('--max_bars', type=int, default=128)
    args = ap.parse_args(argv)
    conf = yaml.load(open(args.configuration), Loader=yaml.FullLoader)
    rep, mode, out_dir = args.representation, args.mode, args.output_dir
    if mode == 'lead_sheet':
        temp, top_p, max_dec_len, emotions = 1.2, 0.97, 512, ['Positive', 'Negative']
    else:
        temp, top_p, max_dec_len, emotions = 1.1, 0.99, 2400, ['Q1', 'Q2', 'Q3', 'Q4']
    print('[nucleus parameters] t = {}, p = {}'.format(temp, top_p))
    os.makedirs(out_dir, exist_ok=True)
    if args.synthetic:
        event2idx, idx2event = synthetic_vocab(args.synthetic, 1)
        vocab_size = args.synthetic
        key_determine = None                         # random weights rarely emit a Key_* first
    else:
        event2idx, idx2event, vocab_size = load_dictionary(conf['data']['vocab_path'].format(rep))   # + PAD (read_vocab)
        key_determine = 'rule'
    tgt_len = conf['model']['decoder']['tgt_len']
    model = build_model(conf, vocab_size, mem_len=tgt_len)
    print('[info] # params:', sum(p.numel() for p in model.parameters() if p.requires_grad))
    if args.inference_params:
        model.load_state_dict(torch.load(args.inference_params, map_location='cpu'))
    model.eval()
    shutil.copy(args.configuration, os.path.join(out_dir, 'config_lead.yaml' if mode == 'lead_sheet' else 'config_full.yaml'))
    gen_times = []
    from ..decode import Stage1Decoder
    dec = Stage1Decoder(model, batch=1, max_len=max_dec_len + 1024)      # one K | V cache + step graphs for every piece
    for piece in range(int(args.n_groups)):
        for emotion in emotions:
            out_name = 'samp_{:02d}_{}'.format(piece, emotion)
            gen_words, t_sec = generate_plain_xl(model, event2idx, idx2event, max_events=max_dec_len, max_bars=args.max_bars,
                                                 primer=['Emotion_{}'.format(emotion)], temp=temp, top_p=top_p,
                                                 representation=rep, key_determine=key_determine, decoder=dec)
            if gen_words is None:
                continue
            events = [idx2event[w] for w in gen_words]
            write_lead_sheet_outputs(out_dir, out_name, events, rep, mode)
            gen_times.append(t_sec)
    if gen_times:
        print('[info] finished generating {} pieces, avg. time: {:.2f} +/- {:.2f} secs.'.format(
            len(gen_times), np.mean(gen_times), np.std(gen_times)))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
