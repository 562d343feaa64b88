"""Stage-1 generation -- the reference's `python3 stage1_compose/inference.py -c CONFIG -r REPR -m MODE [-i PARAMS]
[-o OUT_DIR] [-n N_GROUPS]` surface (inference.py:86-298): builds the model with mem_len = tgt_len, generates
one lead sheet per valence (Positive / Negative) per group with t = 1.2, p = 0.97 (lead_sheet) and writes the
`samp_XX_<Emotion>[_roman].txt` event files stage 2 consumes.  Event->MIDI conversion and the functional
(roman -> absolute) conversion are the reference's own host-side code when its tree is importable."""
import argparse
import os
import shutil
import numpy as np
import torch
import yaml

from ..generate import generate_plain_xl
from ..synth import synthetic_vocab
from . import common
from .stage1_train import build_model


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
        ut = common.reference_module('stage1_compose', 'utils')
        event2idx, idx2event = ut.pickle_load(conf['data']['vocab_path'].format(rep))
        vocab_size = len(event2idx) + 1              # + PAD, as stage1 inference.py read_vocab
        key_determine = 'rule'
    tgt_len = conf['model']['decoder']['tgt_len']
    model = build_model(conf, vocab_size, mem_len=tgt_len)
    print('[info] # params:', sum(p.numel() for p in model.parameters() if p.requires_grad))
    if args.inference_params:
        model.load_state_dict(torch.load(args.inference_params, map_location='cpu'))
    model.eval()
    shutil.copy(args.configuration, os.path.join(out_dir, 'config_lead.yaml' if mode == 'lead_sheet' else 'config_full.yaml'))
    gen_times = []
    for piece in range(int(args.n_groups)):
        for emotion in emotions:
            out_name = 'samp_{:02d}_{}'.format(piece, emotion)
            gen_words, t_sec = generate_plain_xl(model, event2idx, idx2event, max_events=max_dec_len, max_bars=args.max_bars,
                                                 primer=['Emotion_{}'.format(emotion)], temp=temp, top_p=top_p,
                                                 representation=rep, key_determine=key_determine)
            if gen_words is None:
                continue
            events = [idx2event[w] for w in gen_words]
            suffix = '_roman.txt' if rep == 'functional' else '.txt'
            with open(os.path.join(out_dir, out_name + suffix), 'w') as f:
                print(*events[1:], sep='\n', file=f)
            gen_times.append(t_sec)
    if gen_times:
        print('[info] finished generating {} pieces, avg. time: {:.2f} +/- {:.2f} secs.'.format(
            len(gen_times), np.mean(gen_times), np.std(gen_times)))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
