"""Stage-2 generation -- the reference's `python3 stage2_accompaniment/inference.py -m MODEL -c CONFIG -r REPR
[-i PARAMS] [-o OUT_DIR] [-p]` surface (inference.py:330-485) over the incremental B200 decode path.

Reads every generated lead sheet (`*_roman.txt` / `*.txt`) in OUT_DIR, decodes one accompaniment per emotion
quadrant (Positive -> Q1,Q4; Negative -> Q2,Q3) with the reference's sampling constants (performer: t=1.1,
p=0.99; gpt2: t=1.2, p=0.97) and writes `<name>_<Q>_full.txt` token-event files and the `.mid` of the Full track
(`data/midi_out.py`, inference.py:462-479).  `--synthetic V` runs on a synthetic vocabulary / random lead sheets / random weights."""
import argparse
import os
import shutil
import time
from itertools import chain
import numpy as np
import torch
import yaml

from ..decode import Stage2Decoder
from ..generate import generate_conditional, generate_conditional_batch
from ..synth import synthetic_vocab, synthetic_lead_sheet
from ..data.formats import load_dictionary, read_lead_sheet, lead_sheet_files, emotions_for
from ..data.midi_out import relative_to_absolute, full_track_bars, events_to_score, write_midi
from .stage2_train import build_model, load_params

MAX_BARS = 128


def write_accompaniment_midi(out_dir, name, key, events, rep, max_bars=MAX_BARS):
    """the Full-track events of the generated bars -> `<name>.mid` (inference.py:172-208,468-479); False when the
    sequence holds no complete note (random weights)"""
    try:
        absolute = relative_to_absolute(key, events) if rep == 'functional' else events
        bars = full_track_bars(absolute)
        score = events_to_score(key, list(chain(*bars[:max_bars])), mode='full')
        write_midi(os.path.join(out_dir, name + '.mid'), score)
        return True
    except (ValueError, AssertionError, KeyError, TypeError) as err:
        print('[info] {}: no MIDI written ({})'.format(name, type(err).__name__))
        return False


def main(argv=None):
    ap = argparse.ArgumentParser(description='')
    req = ap.add_argument_group('required arguments')
    req.add_argument('-m', '--model_type', choices=['performer', 'gpt2'], required=True, help='model backbone')
    req.add_argument('-c', '--configuration', required=True, help='configurations of training')
    req.add_argument('-r', '--representation', choices=['remi', 'functional'], required=True)
    ap.add_argument('-i', '--inference_params', default=None, help='inference parameters')
    ap.add_argument('-o', '--output_dir', default='generation/emopia_functional_two', help='output directory')
    ap.add_argument('-p', '--play_midi', default=False, action='store_true')
    ap.add_argument('--synthetic', type=int, default=0, help='vocabulary size of a synthetic run (no dataset needed)')
    ap.add_argument('--max_bars', type=int, default=MAX_BARS)
    ap.add_argument('--lockstep', default=False, action='store_true',
                    help='decode the quadrants of a lead sheet together (one batched model step per iteration)')
    args = ap.parse_args(argv)
    conf = yaml.load(open(args.configuration), Loader=yaml.FullLoader)
    tc, mc = conf['training'], conf['model']
    gpuid = tc['gpuid']
    torch.cuda.set_device(gpuid)
    rep, out_dir = args.representation, args.output_dir
    os.makedirs(out_dir, exist_ok=True)

    if args.synthetic:
        event2idx, idx2event = synthetic_vocab(args.synthetic, 2)
        vocab_size = args.synthetic
        for i, emo in enumerate(('Positive', 'Negative')):        # two synthetic lead sheets -> the 4Q run
            path = os.path.join(out_dir, 'samp_%02d_%s%s.txt' % (i, emo, '_roman' if rep == 'functional' else ''))
            if not os.path.exists(path):
                bars = synthetic_lead_sheet(event2idx, 4, i)
                with open(path, 'w') as f:
                    print('Key_C', *[idx2event[t] for t in chain(*bars)], sep='\n', file=f)
    else:
        # the reference builds the validation dataset only to read the vocabulary off it (inference.py:374-381)
        event2idx, idx2event, vocab_size = load_dictionary(conf['data_loader']['vocab_path'].format(rep))

    model = build_model(args.model_type, vocab_size, mc, gpuid)
    temp, top_p = (1.1, 0.99) if args.model_type == "performer" else (1.2, 0.97)
    print(f"[info] temp = {temp} | top_p = {top_p}")
    if args.inference_params:
        load_params(model, args.inference_params)
    model.eval()
    print('[info] model loaded')
    shutil.copy(args.configuration, os.path.join(out_dir, 'config_full.yaml'))
    files = lead_sheet_files(out_dir, rep)
    print('[# pieces]', len(files))
    dec = Stage2Decoder(model, batch=1)
    n_tok, t0 = 0, time.time()
    decs = {}
    for file in files:
        out_name = '_'.join(os.path.basename(file).split('_')[:2])
        if getattr(args, 'lockstep', False):
            # the quadrants of one lead sheet decoded in lockstep (generate_conditional_batch): same rules per sequence, one
            # batched model step per iteration; the numpy RNG is consumed in a different order than sequential calls
            todo = [e for e in emotions_for(file) if not os.path.exists(os.path.join(out_dir, out_name + '_' + e + '_full.txt'))]
            if todo:
                key, lead = read_lead_sheet(file, event2idx)
                if not lead:              # (the reference indexes bar 0 and dies with an IndexError here, inference.py:237)
                    print('[info] {} holds no complete bar, skipping ...'.format(file))
                    continue
                primers = [[event2idx['Emotion_{}'.format(e)]] + ([event2idx[key]] if rep == 'functional' else []) +
                           [event2idx['Tempo_{}'.format(110)]] for e in todo]
                if len(todo) not in decs:
                    decs[len(todo)] = Stage2Decoder(model, batch=len(todo))
                outs = generate_conditional_batch(model, event2idx, idx2event, [lead] * len(todo), primers, [temp] * len(todo),
                                                  top_p=top_p, max_bars=args.max_bars, decoder=decs[len(todo)])
                for e, generated in zip(todo, outs):
                    n_tok += len(generated)
                    events = [idx2event[w] for w in generated]
                    with open(os.path.join(out_dir, out_name + '_' + e + '_full.txt'), 'w') as f:
                        print(*events, sep='\n', file=f)
                    write_accompaniment_midi(out_dir, out_name + '_' + e + '_full', key, events, rep, args.max_bars)
            continue
        for e in emotions_for(file):
            out_txt = os.path.join(out_dir, out_name + '_' + e + '_full.txt')
            if os.path.exists(out_txt):
                print('[info] {} exists, skipping ...'.format(out_txt))
                continue
            key, lead = read_lead_sheet(file, event2idx)
            if not lead:
                print('[info] {} holds no complete bar, skipping ...'.format(file))
                break
            primer = [event2idx['Emotion_{}'.format(e)]] + ([event2idx[key]] if rep == 'functional' else []) + \
                     [event2idx['Tempo_{}'.format(110)]]
            generated = generate_conditional(model, event2idx, idx2event, lead, primer=primer, max_bars=args.max_bars,
                                             temp=temp, top_p=top_p, inadmissibles=None, model_type=args.model_type,
                                             decoder=dec)
            n_tok += len(generated)
            events = [idx2event[w] for w in generated]
            with open(out_txt, 'w') as f:
                print(*events, sep='\n', file=f)
            write_accompaniment_midi(out_dir, out_name + '_' + e + '_full', key, events, rep, args.max_bars)
    dt = time.time() - t0
    print('[info] %d events in %.2f s (%.1f events/s incl. grammar rejections and host loop)' % (n_tok, dt, n_tok / max(dt, 1e-9)))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
