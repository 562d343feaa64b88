"""Stage-1 training -- the reference's `python3 stage1_compose/train.py -c CONFIG -r {remi,functional}`
surface (train.py:19-106,191-359: YAML schema, per-segment loop with max_n_seg = 1, clip 0.5 + Adam, linear
warm-up + cosine decay, log / checkpoint names) over the B200 hot path.  `[T, B]` token layout as in the
reference; under torchrun the batches are rank-strided with one gradient all-reduce per step."""
import argparse
import os
import shutil
import time
import numpy as np
import torch
import yaml

from .. import dp
from ..optim import FusedAdam, WarmupCosine
from ..stage1 import PlainTransformer
from ..synth import synthetic_batch
from . import common


def build_model(conf, vocab_size, mem_len=None, compute_dtype=torch.bfloat16):
    mc = conf['model']
    dc = mc['decoder']
    return PlainTransformer(mc['d_word_embed'], vocab_size, dc['n_layer'], dc['n_head'], dc['d_model'], dc['d_ff'],
                            dc['mem_len'] if mem_len is None else mem_len, dc['tgt_len'], dec_dropout=dc['dropout'],
                            pre_lnorm=mc['pre_lnorm'], compute_dtype=compute_dtype).cuda()


class SyntheticStage1:
    def __init__(self, V, B, T, n, seed):
        self.batches = []
        for i in range(n):
            tok, _, _ = synthetic_batch(V, B, T, seed + i)
            tgt = torch.roll(tok, -1, 1)
            tgt[:, -1] = V - 1
            self.batches.append({'id': torch.arange(B), 'n_seg': [1] * B, 'dec_inp_0': tok, 'dec_tgt_0': tgt,
                                 'dec_seg_len_0': torch.full((B,), T)})

    def __iter__(self):
        return iter(self.batches)


def main(argv=None):
    ap = argparse.ArgumentParser(description='')
    req = ap.add_argument_group('required arguments')
    req.add_argument('-c', '--configuration', required=True, help='configurations of training')
    req.add_argument('-r', '--representation', choices=['remi', 'functional'], required=True)
    ap.add_argument('--max_steps', type=int, default=None)
    args = ap.parse_args(argv)
    conf = yaml.load(open(args.configuration), Loader=yaml.FullLoader)
    rep = args.representation
    tc, dc = conf['training'], conf['data']
    rank, local, world = dp.init_from_env()
    torch.cuda.set_device(local if world > 1 else 0)
    ckpt_dir = conf['output']['ckpt_dir'].format(rep)
    params_dir, optim_dir = os.path.join(ckpt_dir, 'params/'), os.path.join(ckpt_dir, 'optim/')

    if dc.get('synthetic'):
        sy = dc['synthetic']
        vocab_size = sy['vocab_size']
        T = conf['model']['decoder']['tgt_len']
        dloader = SyntheticStage1(vocab_size, dc['batch_size'], T, sy['n_train_batches'], 0)
        vloader = SyntheticStage1(vocab_size, dc['batch_size'], T, sy['n_val_batches'], 10 ** 6)
    elif dc.get('gpu_token_store'):
        # pieces tokenised once into HBM, a batch = one kernel launch (data/token_store.py: Stage1TokenStore); same batch
        # dicts as the reference Dataset + collate_fn, minus the encoder features the loop never reads
        import pickle
        from ..data import Stage1TokenStore
        ddir, vfile = dc['data_dir'].format(rep), dc['vocab_path'].format(rep)
        mk = lambda split: Stage1TokenStore.from_files(
            [os.path.join(ddir, p_) for p_ in pickle.load(open(split, 'rb')) if os.path.exists(os.path.join(ddir, p_))], vfile,
            model_dec_seqlen=conf['model']['decoder']['tgt_len'], device='cuda')
        dset, vset = mk(dc['train_split']), mk(dc['val_split'])
        vocab_size = dset.vocab_size

        dloader = common.StoreEpochs(dset, dc['batch_size'], True, rank, world, seed=dc.get('seed', 0))
        vloader = common.StoreEpochs(vset, dc['batch_size'], False)                # rank 0 validates alone
    else:
        from torch.utils.data import DataLoader
        dl = common.reference_module('stage1_compose', 'dataloader')
        ut = common.reference_module('stage1_compose', 'utils')
        mk = lambda split: dl.SkylineFullSongTransformerDataset(
            dc['data_dir'].format(rep), dc['vocab_path'].format(rep), pieces=ut.pickle_load(split),
            do_augment=False, model_dec_seqlen=conf['model']['decoder']['tgt_len'], max_n_seg=dc['max_n_seg'],
            max_pitch=108, min_pitch=48, convert_dict_event=True)
        dset, vset = mk(dc['train_split']), mk(dc['val_split'])
        vocab_size = dset.vocab_size
        dloader = DataLoader(dset, batch_size=dc['batch_size'], shuffle=True, num_workers=24, collate_fn=dset.collate_fn,
                             generator=common.shared_generator(dc.get('seed', 0)))   # same permutation on every rank
        vloader = DataLoader(vset, batch_size=dc['batch_size'], num_workers=8, collate_fn=vset.collate_fn)

    torch.manual_seed(0)
    model = build_model(conf, vocab_size)
    if conf.get('pretrained_param_path'):
        model.load_state_dict(torch.load(conf['pretrained_param_path'], map_location='cpu'))      # strict (train.py:213-228)
    model.train()
    sync = dp.GradSync(model)
    sync.broadcast_params()
    print('[info] # params:', sum(p.numel() for p in model.parameters() if p.requires_grad))
    opt = FusedAdam(model, lr=tc['max_lr'], max_grad_norm=0.5, grad_sync=sync)
    sched = WarmupCosine(opt, tc['max_lr'], tc['min_lr'], tc['warmup_steps'], tc['lr_decay_steps'])
    if conf.get('pretrained_optim_path'):
        opt.load_state_dict(torch.load(conf['pretrained_optim_path'], map_location='cpu'))
    if rank == 0:
        for d_ in (ckpt_dir, params_dir, optim_dir):
            os.makedirs(d_, exist_ok=True)
        shutil.copy(args.configuration, os.path.join(ckpt_dir, 'config.yaml'))

    steps = tc.get('trained_steps', 0)
    for ep in range(tc.get('trained_epochs', 0) + 1, tc['max_epoch'] + 1):
        model.train()
        st = time.time()
        stats = torch.zeros(3, device='cuda')
        for batch in common.rank_strided(dloader, rank, world):
            for seg_i in range(max(batch['n_seg'])):                                      # = 1 in every config
                x = batch['dec_inp_%d' % seg_i].permute(1, 0).cuda(non_blocking=True)       # [T, B]
                tgt = batch['dec_tgt_%d' % seg_i].permute(1, 0).cuda(non_blocking=True)
                steps += 1
                acc = model.train_step(x, tgt, count_allreduce=sync.count_allreduce)
                opt.step()
                sched.update(steps)
                stats += sync.allreduce_stats(acc) if world > 1 else acc
            if rank == 0 and steps % tc['log_interval'] == 0:
                s = stats.tolist()
                with open(os.path.join(ckpt_dir, 'log.txt'), 'a') as f:
                    f.write('{:<4} {:<8} {:<12} {:<10} {:<12}\n'.format(ep, steps, round(s[1] / max(s[0], 1), 5),
                                                                        round(s[2] / max(s[0], 1), 4), round(time.time() - st, 2)))
            if args.max_steps and steps >= args.max_steps:
                break
        s = stats.tolist()
        loss = s[1] / max(s[0], 1.0)
        if rank == 0:
            print('[epoch {:03d}] loss = {:.4f}, total_acc = {:.4f}, time = {:.2f} secs'.format(ep, loss, s[2] / max(s[0], 1), time.time() - st))
            if ep % conf['output']['ckpt_interval'] == 0:
                torch.save(model.state_dict(), os.path.join(params_dir, 'ep{:03d}_loss{:.3f}_params.pt'.format(ep, loss)))
                torch.save(opt.state_dict(), os.path.join(optim_dir, 'ep{:03d}_loss{:.3f}_optim.pt'.format(ep, loss)))
            if ep % tc.get('val_interval', 1) == 0:
                model.eval()
                vl = []
                with torch.no_grad():
                    for batch in vloader:
                        x = batch['dec_inp_0'].permute(1, 0).cuda()
                        tgt = batch['dec_tgt_0'].permute(1, 0).cuda()
                        logits, _ = model(x, tuple())
                        vl.append(float(model.compute_loss(logits, tgt)['ce_loss']))
                with open(os.path.join(ckpt_dir, 'valloss.txt'), 'a') as f:
                    f.write('ep{:03d} | loss: {:.3f} | valloss: {:.3f} (±{:.3f})\n'.format(ep, loss, np.mean(vl), np.std(vl)))
        if args.max_steps and steps >= args.max_steps:
            break
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
